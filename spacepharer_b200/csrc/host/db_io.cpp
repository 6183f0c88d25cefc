// db_io.cpp -- see db_io.h
#include "db_io.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <stdexcept>

namespace spbhost {

static bool file_exists(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 && S_ISREG(st.st_mode);
}

static void append_file(const std::string &p, std::vector<char> &out) {
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
    const std::string idx = name + ".index";
    FILE *f = fopen(idx.c_str(), "r");
    if (!f) throw std::runtime_error("Could not open index file " + idx + "!");
    char line[256];
    while (fgets(line, sizeof(line), f)) {
        char *p = line, *e = nullptr;
        unsigned long long k = strtoull(p, &e, 10);
        if (e == p) continue;
        p = e;
        unsigned long long o = strtoull(p, &e, 10);
        if (e == p) throw std::runtime_error("Malformed index line in " + idx);
        p = e;
        unsigned long long l = strtoull(p, &e, 10);
        if (e == p) throw std::runtime_error("Malformed index line in " + idx);
        if (o + l > dataSize_) throw std::runtime_error("Index entry beyond the end of the data file in " + idx);
        entries_.push_back(DBEntry{(uint32_t)k, (uint64_t)o, (uint64_t)l});
    }
    fclose(f);
    std::stable_sort(entries_.begin(), entries_.end(), [](const DBEntry &a, const DBEntry &b) { return a.key < b.key; });
    linear_.resize(entries_.size());
    for (size_t i = 0; i < entries_.size(); i++) linear_[i] = (uint32_t)i;
    std::stable_sort(linear_.begin(), linear_.end(), [&](uint32_t a, uint32_t b) { return entries_[a].offset < entries_[b].offset; });
    linearRank_.resize(entries_.size());
    for (size_t i = 0; i < linear_.size(); i++) linearRank_[linear_[i]] = (uint32_t)i;
    dbtype_ = 0;
    const std::string tp = name + ".dbtype";
    FILE *t = fopen(tp.c_str(), "rb");
    if (t) {
        unsigned char b[4] = {0, 0, 0, 0};
        if (fread(b, 1, 4, t) == 4) dbtype_ = (int)(b[0] | (b[1] << 8) | (b[2] << 16) | ((unsigned)b[3] << 24));
        fclose(t);
    }
    if (dbtype_ & (1 << 31)) throw std::runtime_error("Compressed databases (--compressed 1) are not supported: " + name);
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

void DBWriter::open(const std::string &name, int dbtype, int threads) {
    name_ = name;
    dbtype_ = dbtype;
    parts_.assign((size_t)(threads > 0 ? threads : 1), Part());
    // a stale .dbtype would make the workflow believe the step is finished
    remove((name + ".dbtype").c_str());
}

void DBWriter::write(int thread, uint32_t key, const char *payload, size_t len) {
    Part &p = parts_[(size_t)thread];
    DBEntry e;
    e.key = key;
    e.offset = p.data.size();
    e.length = len + 1;
    p.data.insert(p.data.end(), payload, payload + len);
    p.data.push_back('\0');
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

void DBWriter::close() {
    FILE *f = fopen(name_.c_str(), "wb");
    if (!f) throw std::runtime_error("Could not open " + name_ + " for writing!");
    std::vector<DBEntry> all;
    uint64_t base = 0;
    bool ok = true;
    for (Part &p : parts_) {
        if (!p.data.empty() && fwrite(p.data.data(), 1, p.data.size(), f) != p.data.size()) ok = false;
        for (DBEntry e : p.idx) {
            e.offset += base;
            all.push_back(e);
        }
        base += p.data.size();
    }
    if (fclose(f) != 0) ok = false;
    if (!ok) {
        remove(name_.c_str());
        throw std::runtime_error("Could not write to " + name_);
    }
    std::stable_sort(all.begin(), all.end(), [](const DBEntry &a, const DBEntry &b) { return a.key < b.key; });
    // index lines "key\toffset\tlength\n", formatted into one buffer (no stdio call per entry)
    std::string ix;
    ix.resize(all.size() * 64 + 64);
    char *w = &ix[0];
    for (const DBEntry &e : all) {
        w = put_u64(w, e.key); *w++ = '\t';
        w = put_u64(w, (unsigned long long)e.offset); *w++ = '\t';
        w = put_u64(w, (unsigned long long)e.length); *w++ = '\n';
    }
    write_all(name_ + ".index", "wb", ix.data(), (size_t)(w - ix.data()));
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
