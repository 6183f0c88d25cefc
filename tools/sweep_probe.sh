#!/bin/bash
# prefilter tuning sweep (development aid): kernel time of the team prefilter for a few (kCapSolo, hitCapSolo) settings
for cfg in "4096 8192" "2048 8192" "8192 8192" "12000 8192" "16384 16384" "32768 32768" "65535 65536"; do
  set -- $cfg
  echo "== kcap $1 hitcap $2"
  SPB_KCAP_SOLO=$1 SPB_HITCAP_SOLO=$2 timeout 300 python tools/perf_probe.py 8000 10000 2>&1 | grep "prefilter wall" | tail -1 | cut -c1-260
done
