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
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
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

void DBWriter::close() {
    FILE *f = fopen(name_.c_str(), "wb");
    if (!f) throw std::runtime_error("Could not open " + name_ + " for writing!");
    std::vector<DBEntry> all;
    uint64_t base = 0;
    for (Part &p : parts_) {
        if (!p.data.empty() && fwrite(p.data.data(), 1, p.data.size(), f) != p.data.size()) {
            fclose(f);
            throw std::runtime_error("Could not write to " + name_);
        }
        for (DBEntry e : p.idx) {
            e.offset += base;
            all.push_back(e);
        }
        base += p.data.size();
    }
    fclose(f);
    std::stable_sort(all.begin(), all.end(), [](const DBEntry &a, const DBEntry &b) { return a.key < b.key; });
    FILE *ix = fopen((name_ + ".index").c_str(), "w");
    if (!ix) throw std::runtime_error("Could not open " + name_ + ".index for writing!");
    for (const DBEntry &e : all) fprintf(ix, "%u\t%llu\t%llu\n", e.key, (unsigned long long)e.offset, (unsigned long long)e.length);
    fclose(ix);
    FILE *t = fopen((name_ + ".dbtype").c_str(), "wb");
    if (!t) throw std::runtime_error("Could not open " + name_ + ".dbtype for writing!");
    unsigned char b[4] = {(unsigned char)(dbtype_ & 0xFF), (unsigned char)((dbtype_ >> 8) & 0xFF), (unsigned char)((dbtype_ >> 16) & 0xFF),
                          (unsigned char)((dbtype_ >> 24) & 0xFF)};
    fwrite(b, 1, 4, t);
    fclose(t);
    parts_.clear();
}

}  // namespace spbhost
