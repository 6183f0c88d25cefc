#!/bin/bash
# Final-TSV parity check (development time; needs the reference build and a reference easy-predict run kept with
# --remove-tmp-files 0): the reference's own post-processing steps (data/predictmatch.sh:127-168) are run on the `aggregate` and
# `match` DBs written by `spb200 predictmatch` on the B200 (tools/tsv_gpu_side.py -> gpurun_out/tsv) and the resulting TSV is
# compared with the reference's result.tsv.
# Usage: tools/tsv_parity_check.sh <reference work dir, e.g. /tmp/nw/c1> [reference binary]
set -eu
W=$1
B=${2:-/tmp/refbuild/src/spacepharer}
OURS=$(dirname "$0")/../gpurun_out/tsv
P=$(dirname "$(find "$W/tmp" -name nucl_result.index | head -1)")
Q=$(dirname "$(find "$W/tmp" -name qsetdb.index | head -1)")/qsetdb
T=$(mktemp -d)
for n in aggregate match; do cp "$OURS/$n" "$OURS/$n.index" "$OURS/$n.dbtype" "$T/"; done
$B truncatebesthits "$Q" "$T/aggregate" "$T/aggregate_truncated" --threads 4 -v 1
$B offsetalignment "${Q}_nucl" "$Q" "$W/targets_nucl" "$W/targets" "$T/aggregate_truncated" "$T/aggregate_offset" --search-type 4 --threads 4 -v 1
$B convertalis "${Q}_nucl" "$W/targets_nucl" "$T/aggregate_offset" "$T/aln" --db-output --search-type 3 \
   --format-output tsetid,query,qset,target,evalue,qstart,qend,qlen,tstart,tend,qaln,taln,empty --threads 4 -v 1
$B mergeresultsbyset "${Q}_nucl_set_to_contig" "$T/aln" "$T/aln_merge" --threads 4 -v 1
$B findpam "$W/targets_nucl" "$T/aln_merge" "$T/aln_merge_pam" --flanking-seq-len 10 --compressed 0 --threads 4 -v 1
$B summarizeresults "$W/targets_nucl" "$T/match" "$T/aln_merge_pam" "$T/result.tsv" --fmt 1 --report-fdr 0 \
   --lca-ranks superkingdom,phylum,class,order,family,genus,species --tax-lineage 0 --compressed 0 --threads 4 -v 1
if cmp -s "$T/result.tsv" "$W/result.tsv"; then
  echo "result.tsv IDENTICAL to the reference's ($(wc -l < "$T/result.tsv") lines, sha256 $(sha256sum < "$T/result.tsv" | cut -c1-16)...)"
else
  echo "result.tsv DIFFERS"; diff "$T/result.tsv" "$W/result.tsv" | head
  exit 1
fi
