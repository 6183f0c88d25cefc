// db_io.cpp -- see db_io.h
#include "db_io.h"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#ifdef _OPENMP
#include <omp.h>
#include <parallel/algorithm>
#endif

#include <algorithm>
#include <stdexcept>

namespace spbhost {

// ---- zstd through the system library (no headers in the image: the four entry points are declared here) ----
namespace {
struct Zstd {
    size_t (*decompress)(void *dst, size_t dstCapacity, const void *src, size_t srcSize) = nullptr;
    size_t (*compress)(void *dst, size_t dstCapacity, const void *src, size_t srcSize, int level) = nullptr;
    size_t (*compressBound)(size_t srcSize) = nullptr;
    unsigned (*isError)(size_t code) = nullptr;
    const char *(*getErrorName)(size_t code) = nullptr;
};

const Zstd &zstd() {
    static const Zstd z = [] {
        Zstd r;
        void *h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libzstd.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) throw std::runtime_error("Compressed databases (--compressed 1) need libzstd.so.1, which could not be loaded");
        r.decompress = (decltype(r.decompress))dlsym(h, "ZSTD_decompress");
        r.compress = (decltype(r.compress))dlsym(h, "ZSTD_compress");
        r.compressBound = (decltype(r.compressBound))dlsym(h, "ZSTD_compressBound");
        r.isError = (decltype(r.isError))dlsym(h, "ZSTD_isError");
        r.getErrorName = (decltype(r.getErrorName))dlsym(h, "ZSTD_getErrorName");
        if (!r.decompress || !r.compress || !r.compressBound || !r.isError || !r.getErrorName)
            throw std::runtime_error("libzstd.so.1 lacks the simple API (ZSTD_compress / ZSTD_decompress)");
        return r;
    }();
    return z;
}

bool g_compressedDefault = false;
}  // namespace

static bool file_exists(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

static void append_file(const std::string &p, RawBuf &out) {
    FILE *f = fopen(p.c_str(), "rb");
    if (!f) throw std::runtime_error("Could not open data file " + p + "!");
    long n = -1;
    if (fseek(f, 0, SEEK_END) == 0) n = ftell(f);
    if (n < 0 || fseek(f, 0, SEEK_SET) != 0) {
        fclose(f);
        throw std::runtime_error("Failed to read data file " + p);
    }
    size_t at = out.size();
    out.resize(at + (size_t)n);
    if (n > 0 && fread(out.data() + at, 1, (size_t)n, f) != (size_t)n) {
        fclose(f);
        throw std::runtime_error("Failed to read data file " + p);
    }
    fclose(f);
}

void DBReader::open(const std::string &name) {
    entries_.clear();
    data_.clear();
    // data: <name>, else <name>.0, <name>.1, ... (DBReader.cpp:105-160)
    if (file_exists(name)) {
        append_file(name, data_);
    } else {
        int n = 0;
        while (file_exists(name + "." + std::to_string(n))) {
            append_file(name + "." + std::to_string(n), data_);
            n++;
        }
        if (n == 0) throw std::runtime_error("Could not open data file " + name + "!");
    }
    dataSize_ = data_.size();
    dbtype_ = 0;
    {
        const std::string tp = name + ".dbtype";
        FILE *t = fopen(tp.c_str(), "rb");
        if (t) {
            unsigned char b[4] = {0, 0, 0, 0};
            if (fread(b, 1, 4, t) == 4) dbtype_ = (int)(b[0] | (b[1] << 8) | (b[2] << 16) | ((unsigned)b[3] << 24));
            fclose(t);
        }
    }
    const bool isCompressed = compressed();      // index lengths are uncompressed lengths then
    // index: the whole text is read once and parsed by all host threads (slices cut at line ends); a 1 M-spacer query DB has
    // 3.8 M lines, and stdio + strtoull per line used to cost more than the GPU search itself
    const std::string idx = name + ".index";
    RawBuf text;
    {
        FILE *f = fopen(idx.c_str(), "r");
        if (!f) throw std::runtime_error("Could not open index file " + idx + "!");
        fclose(f);
        append_file(idx, text);
    }
    const size_t tn = text.size();
    int nth = 1;
#ifdef _OPENMP
    nth = std::max(1, omp_get_max_threads());
#endif
    if (tn < (1u << 20)) nth = 1;
    // two passes over slices cut at line ends: count the lines, then parse straight into the final table (sized once, not
    // zeroed: every thread touches its own stretch first -- the page faults of a 3.8 M-entry table were most of the parse time)
    std::vector<size_t> cut((size_t)nth + 1, 0), lines((size_t)nth + 1, 0);
    for (int t = 0; t <= nth; t++) {
        size_t a = tn * (size_t)t / (size_t)nth;
        if (t > 0 && t < nth) { while (a < tn && text[a - 1] != '\n') a++; }
        cut[(size_t)t] = (t == nth) ? tn : a;
    }
#pragma omp parallel for num_threads(nth) schedule(static, 1)
    for (int t = 0; t < nth; t++) {
        size_t n = 0;
        const char *p = text.data() + cut[(size_t)t], *const end = text.data() + cut[(size_t)t + 1];
        while (p < end) {
            const char *le = (const char *)memchr(p, '\n', (size_t)(end - p));
            n++;
            if (!le) break;
            p = le + 1;
        }
        lines[(size_t)t + 1] = n;
    }
    for (int t = 0; t < nth; t++) lines[(size_t)t + 1] += lines[(size_t)t];
    entries_.resize(lines[(size_t)nth]);
    std::vector<size_t> used((size_t)nth, 0);
    std::vector<int> bad((size_t)nth, 0);
#pragma omp parallel for num_threads(nth) schedule(static, 1)
    for (int t = 0; t < nth; t++) {
        DBEntry *out = entries_.data() + lines[(size_t)t];
        size_t n = 0;
        const char *p = text.data() + cut[(size_t)t], *const end = text.data() + cut[(size_t)t + 1];
        while (p < end) {
            const char *le = (const char *)memchr(p, '\n', (size_t)(end - p));
            if (!le) le = end;
            unsigned long long v[3] = {0, 0, 0};
            int nf = 0;
            const char *q = p;
            while (nf < 3) {
                while (q < le && (*q == ' ' || *q == '\t' || *q == '\r')) q++;
                if (q >= le || *q < '0' || *q > '9') break;
                unsigned long long x = 0;
                while (q < le && *q >= '0' && *q <= '9') x = x * 10 + (unsigned long long)(*q++ - '0');
                v[nf++] = x;
            }
            if (nf == 3) {
                if (isCompressed ? v[1] + 5 > dataSize_ : v[1] + v[2] > dataSize_) bad[(size_t)t] = 2;
                out[n++] = DBEntry((uint32_t)v[0], (uint64_t)v[1], (uint64_t)v[2]);
            } else if (nf != 0) {
                bad[(size_t)t] = 1;
            }
            p = le + 1;
        }
        used[(size_t)t] = n;
    }
    for (int t = 0; t < nth; t++) {
        if (bad[(size_t)t] == 1) throw std::runtime_error("Malformed index line in " + idx);
        if (bad[(size_t)t] == 2) throw std::runtime_error("Index entry beyond the end of the data file in " + idx);
    }
    {   // blank lines leave holes behind a slice: close them (none in a DB any writer produced)
        size_t w = 0;
        for (int t = 0; t < nth; t++) {
            const size_t from = lines[(size_t)t], n = used[(size_t)t];
            if (w != from && n) memmove((void *)(entries_.data() + w), (const void *)(entries_.data() + from), n * sizeof(DBEntry));
            w += n;
        }
        entries_.resize(w);
    }
    // DBs written by the reference's (and our) writers are already in key order with ascending offsets: both sorts are skipped then
    const long ne = (long)entries_.size();
    int keyBad = 0, offBad = 0;
#pragma omp parallel for num_threads(nth) schedule(static) reduction(| : keyBad)
    for (long i = 1; i < ne; i++) keyBad |= entries_[(size_t)i - 1].key > entries_[(size_t)i].key ? 1 : 0;
    if (keyBad)
        std::stable_sort(entries_.begin(), entries_.end(), [](const DBEntry &a, const DBEntry &b) { return a.key < b.key; });
    linear_.resize(entries_.size());
#pragma omp parallel for num_threads(nth) schedule(static) reduction(| : offBad)
    for (long i = 0; i < ne; i++) {
        linear_[(size_t)i] = (uint32_t)i;
        if (i > 0) offBad |= entries_[(size_t)i - 1].offset >= entries_[(size_t)i].offset ? 1 : 0;
    }
    if (offBad)
        std::stable_sort(linear_.begin(), linear_.end(), [&](uint32_t a, uint32_t b) { return entries_[a].offset < entries_[b].offset; });
    linearRank_.resize(entries_.size());
#pragma omp parallel for num_threads(nth) schedule(static)
    for (long i = 0; i < ne; i++) linearRank_[linear_[(size_t)i]] = (uint32_t)i;
    if (isCompressed) inflate(name);
}

// DBReader::getDataCompressed (DBReader.cpp:512-539) for every entry at once: the DB is rebuilt in memory in the uncompressed
// layout, entries in their old (offset) order, so that ids, lengths and getAminoAcidDBSize (the sum of the index lengths,
// DBReader.cpp:774-831) are what the reference sees through its per-entry decompression.
void DBReader::inflate(const std::string &name) {
    const Zstd &z = zstd();
    const size_t n = entries_.size();
    std::vector<uint64_t> at(n + 1, 0);
    for (size_t id = 0; id < n; id++) {
        const DBEntry &e = entries_[linear_[id]];
        if (e.length == 0) throw std::runtime_error("Malformed index line in " + name + ".index");
        at[id + 1] = at[id] + e.length;
    }
    RawBuf out((size_t)at[n]);
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(max : bad)
    for (long id = 0; id < (long)n; id++) {
        DBEntry &e = entries_[linear_[(size_t)id]];
        const char *src = data_.data() + e.offset;
        unsigned int cSize;
        memcpy(&cSize, src, 4);
        const uint64_t want = e.length - 1;
        char *dst = out.data() + at[(size_t)id];
        if (e.offset + 4 + (uint64_t)cSize + 1 > data_.size()) { bad = std::max(bad, 2); continue; }
        if (src[4 + cSize] == 0) {                 // one zstd frame
            const size_t got = z.decompress(dst, (size_t)want, src + 4, cSize);
            if (z.isError(got) || got != want) { bad = std::max(bad, 1); continue; }
        } else {                                   // stored (terminator 0xFF)
            if (cSize != want) { bad = std::max(bad, 1); continue; }
            memcpy(dst, src + 4, cSize);
        }
        dst[want] = '\0';
    }
    if (bad == 2) throw std::runtime_error("Index entry beyond the end of the data file in " + name + ".index");
    if (bad) throw std::runtime_error("Could not decompress an entry of " + name + " (ZSTD_decompress)");
    for (size_t id = 0; id < n; id++) entries_[linear_[id]].offset = at[id];
    data_.swap(out);
    dataSize_ = data_.size();
}

size_t DBReader::findKey(uint32_t key) const {
    auto it = std::lower_bound(entries_.begin(), entries_.end(), key, [](const DBEntry &e, uint32_t k) { return e.key < k; });
    if (it == entries_.end() || it->key != key) return SIZE_MAX;
    return (size_t)(it - entries_.begin());
}

size_t DBReader::linearIdOfKey(uint32_t key) const {
    size_t i = findKey(key);
    return i == SIZE_MAX ? SIZE_MAX : (size_t)linearRank_[i];
}

DBWriter::~DBWriter() { if (fd_ >= 0) ::close(fd_); }

void DBWriter::setCompressedDefault(bool on) { g_compressedDefault = on; }

void DBWriter::open(const std::string &name, int dbtype, int threads) {
    name_ = name;
    dbtype_ = dbtype & 0x7FFFFFFF;
    compressed_ = g_compressedDefault;
    if (compressed_) {
        zstd();                                   // fails here, before anything is written, when the library is missing
        dbtype_ |= (int)(1u << 31);
    }
    parts_.assign((size_t)(threads > 0 ? threads : 1), Part());
    fileEnd_ = 0;
    if (fd_ >= 0) ::close(fd_);
    fd_ = -1;
    // a stale .dbtype would make the workflow believe the step is finished
    remove((name + ".dbtype").c_str());
}

// entries of a part are offsets into its unflushed buffer until flush() re-bases them
void DBWriter::write(int thread, uint32_t key, const char *payload, size_t len) {
    Part &p = parts_[(size_t)thread];
    DBEntry e;
    e.key = key;
    e.offset = p.data.size();
    e.length = len + 1;
    if (!compressed_) {
        p.data.insert(p.data.end(), payload, payload + len);
        p.data.push_back('\0');
    } else if (len < 60) {
        // DBWriter::writeAdd / writeEnd (DBWriter.cpp:279-369): short payloads are stored, terminator 0xFF
        const unsigned int sz = (unsigned int)len;
        const char *b = reinterpret_cast<const char *>(&sz);
        p.data.insert(p.data.end(), b, b + 4);
        p.data.insert(p.data.end(), payload, payload + len);
        p.data.push_back((char)0xFF);
    } else {
        // one zstd frame at the reference's level 3, terminator 0x00; the index keeps the uncompressed length
        const Zstd &z = zstd();
        const size_t at = p.data.size(), cap = z.compressBound(len);
        p.data.resize(at + 4 + cap + 1);
        const size_t got = z.compress(p.data.data() + at + 4, cap, payload, len, 3);
        if (z.isError(got)) throw std::runtime_error(std::string("ZSTD_compress: ") + z.getErrorName(got));
        const unsigned int sz = (unsigned int)got;
        memcpy(p.data.data() + at, &sz, 4);
        p.data[at + 4 + got] = '\0';
        p.data.resize(at + 4 + got + 1);
    }
    p.idx.push_back(e);
}

// Every write is checked: a failed or short write (ENOSPC, I/O error) must not leave a DB behind that looks finished -- the
// workflow scripts take <name>.dbtype as the "step done" marker (blastp.sh:31,39), so it is written last, via a temporary file
// and rename().
static void write_all(const std::string &path, const char *mode, const void *data, size_t n) {
    FILE *f = fopen(path.c_str(), mode);
    if (!f) throw std::runtime_error("Could not open " + path + " for writing!");
    const bool ok = (n == 0 || fwrite(data, 1, n, f) == n);
    const bool closed = (fclose(f) == 0);
    if (!ok || !closed) {
        remove(path.c_str());
        throw std::runtime_error("Could not write to " + path);
    }
}

static inline char *put_u64(char *p, unsigned long long v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

// Data written so far by parts [first, first + count) goes to the file now (called between batches, e.g. from the formatter
// thread of a GPU while the next batch is on the device): every part gets the next free stretch of the file, the parts are
// written with one pwrite stream each, and their buffers are emptied.  The entries keep their final offsets in the index.
void DBWriter::flush(int first, int count) {
    std::lock_guard<std::mutex> lock(mu_);
    if (fd_ < 0) {
        fd_ = ::open(name_.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (fd_ < 0) throw std::runtime_error("Could not open " + name_ + " for writing!");
    }
    const int last = std::min<int>((int)parts_.size(), first + count);
    std::vector<uint64_t> base((size_t)std::max(0, last - first) + 1, fileEnd_);
    for (int i = first; i < last; i++) base[(size_t)(i - first) + 1] = base[(size_t)(i - first)] + parts_[(size_t)i].data.size();
    bool ok = true;
#pragma omp parallel for schedule(dynamic, 1) reduction(&& : ok)
    for (int i = first; i < last; i++) {
        Part &pt = parts_[(size_t)i];
        const std::vector<char> &d = pt.data;
        const uint64_t at = base[(size_t)(i - first)];
        size_t done = 0;
        while (done < d.size()) {
            const ssize_t w = pwrite(fd_, d.data() + done, std::min<size_t>(d.size() - done, (size_t)1 << 28), (off_t)(at + done));
            if (w <= 0) { ok = false; break; }
            done += (size_t)w;
        }
        for (size_t k = pt.flushed; k < pt.idx.size(); k++) pt.idx[k].offset += at;
        pt.flushed = pt.idx.size();
        pt.data.clear();
    }
    fileEnd_ = base.back();
    if (!ok) {
        ::close(fd_);
        fd_ = -1;
        remove(name_.c_str());
        throw std::runtime_error("Could not write to " + name_);
    }
}

void DBWriter::close() {
    flush(0, (int)parts_.size());
    const size_t np = parts_.size();
    size_t nEntries = 0;
    for (size_t i = 0; i < np; i++) nEntries += parts_[i].idx.size();
    {
        std::lock_guard<std::mutex> lock(mu_);
        const bool ok = (::close(fd_) == 0);
        fd_ = -1;
        if (!ok) {
            remove(name_.c_str());
            throw std::runtime_error("Could not write to " + name_);
        }
    }
    std::vector<DBEntry> all;
    all.reserve(nEntries);
    for (size_t i = 0; i < np; i++) all.insert(all.end(), parts_[i].idx.begin(), parts_[i].idx.end());
    bool sorted = true;
    for (size_t i = 1; i < all.size() && sorted; i++) sorted = all[i - 1].key <= all[i].key;
    if (!sorted) {
#ifdef _OPENMP
        __gnu_parallel::stable_sort(all.begin(), all.end(), [](const DBEntry &a, const DBEntry &b) { return a.key < b.key; });
#else
        std::stable_sort(all.begin(), all.end(), [](const DBEntry &a, const DBEntry &b) { return a.key < b.key; });
#endif
    }
    // index lines "key\toffset\tlength\n": slices formatted by all threads, written back to back
    int nth = 1;
#ifdef _OPENMP
    nth = std::max(1, omp_get_max_threads());
#endif
    if (all.size() < 65536) nth = 1;
    std::vector<std::string> ixs((size_t)nth);
#pragma omp parallel for num_threads(nth) schedule(static, 1)
    for (int t = 0; t < nth; t++) {
        const size_t a = all.size() * (size_t)t / (size_t)nth, b = all.size() * (size_t)(t + 1) / (size_t)nth;
        std::string &ix = ixs[(size_t)t];
        ix.resize((b - a) * 64 + 64);
        char *w = &ix[0];
        for (size_t i = a; i < b; i++) {
            const DBEntry &e = all[i];
            w = put_u64(w, e.key); *w++ = '\t';
            w = put_u64(w, (unsigned long long)e.offset); *w++ = '\t';
            w = put_u64(w, (unsigned long long)e.length); *w++ = '\n';
        }
        ix.resize((size_t)(w - ix.data()));
    }
    {
        const std::string path = name_ + ".index";
        FILE *f = fopen(path.c_str(), "wb");
        if (!f) throw std::runtime_error("Could not open " + path + " for writing!");
        bool wok = true;
        for (const std::string &ix : ixs)
            if (!ix.empty() && fwrite(ix.data(), 1, ix.size(), f) != ix.size()) wok = false;
        if (fclose(f) != 0) wok = false;
        if (!wok) {
            remove(path.c_str());
            throw std::runtime_error("Could not write to " + path);
        }
    }
    const unsigned char b[4] = {(unsigned char)(dbtype_ & 0xFF), (unsigned char)((dbtype_ >> 8) & 0xFF), (unsigned char)((dbtype_ >> 16) & 0xFF),
                                (unsigned char)((dbtype_ >> 24) & 0xFF)};
    const std::string tmp = name_ + ".dbtype.tmp";
    write_all(tmp, "wb", b, 4);
    if (rename(tmp.c_str(), (name_ + ".dbtype").c_str()) != 0) {
        remove(tmp.c_str());
        throw std::runtime_error("Could not write to " + name_ + ".dbtype");
    }
    parts_.clear();
}

}  // namespace spbhost
