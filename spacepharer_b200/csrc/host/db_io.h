// db_io.h -- MMseqs2 on-disk DB reader / writer of the spb200 host binary.
//
// Byte-compatible with the reference's DBReader / DBWriter for the files the search path touches
// (M/commons/DBReader.h:57-99, DBReader.cpp:105-368,541-550; M/commons/DBWriter.cpp:216-440):
//   <name>         concatenated entries, each NUL-terminated (sequence entries are "residues\n\0")
//   <name>.index   text lines  key \t offset \t length   (length includes the NUL; any line order)
//   <name>.dbtype  4-byte little-endian int (0 amino acids, 5 alignment results, 7 prefilter results)
// A multi-threaded reference writer may leave <name>.0 .. <name>.N-1 whose offsets live in the concatenated space.
//
// Compressed DBs (`--compressed 1`, bit 31 of the dbtype; DBReader.cpp:191,512-539, DBWriter.cpp:254-397): an entry is
//   u32 size | size bytes | terminator   with terminator 0x00 = one zstd frame, 0xFF = stored as is (payloads below 60 bytes);
// its index length stays the uncompressed length + 1.  The reader inflates such a DB once, in parallel, into the uncompressed
// layout (same entry order), so everything behind it is unchanged; the writer deflates entry by entry in the calling thread.
// zstd comes from the system's libzstd.so.1 (dlopen; the image has the library but no headers) -- without it compressed DBs
// fail loudly.
#pragma once

#include <stdint.h>

#include <memory>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

namespace spbhost {

enum { DBTYPE_AMINO_ACIDS = 0, DBTYPE_ALIGNMENT_RES = 5, DBTYPE_PREFILTER_RES = 7 };

// char buffer whose resize() does not zero the new bytes (they are overwritten by a file read or by the inflater right away; the
// zero fill of a few hundred MB was a measurable part of opening a DB)
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { typedef NoInitAlloc<U> other; };
    NoInitAlloc() {}
    template <class U> NoInitAlloc(const NoInitAlloc<U> &) {}
    template <class U> void construct(U *p) { ::new ((void *)p) U; }
    template <class U, class... A> void construct(U *p, A &&... a) { ::new ((void *)p) U(std::forward<A>(a)...); }
};
typedef std::vector<char, NoInitAlloc<char>> RawBuf;

struct DBEntry {
    uint32_t key;
    uint64_t offset;
    uint64_t length;   // incl. the terminating NUL
    DBEntry() {}       // (not zeroed: the reader sizes its table once and fills it from all threads -- first touch in parallel)
    DBEntry(uint32_t k, uint64_t o, uint64_t l) : key(k), offset(o), length(l) {}
};

class DBReader {
public:
    // throws std::runtime_error with the reference's wording on missing files
    void open(const std::string &name);
    size_t size() const { return entries_.size(); }
    int dbtype() const { return dbtype_ & 0x7FFFFFFF; }
    bool compressed() const { return (dbtype_ & (1u << 31)) != 0; }
    // entries sorted by key (DBReader::NOSORT id order after sortIndex) ...
    const DBEntry &byKey(size_t i) const { return entries_[i]; }
    // ... and the LINEAR_ACCESS order (by data offset, DBReader.cpp:339-368): ids of the prefilter index
    const DBEntry &linear(size_t id) const { return entries_[linear_[id]]; }
    // position (in key order) of a key, or SIZE_MAX
    size_t findKey(uint32_t key) const;
    // id in LINEAR_ACCESS order of a key, or SIZE_MAX
    size_t linearIdOfKey(uint32_t key) const;
    const char *data(const DBEntry &e) const { return data_.data() + e.offset; }
    // DBReader::getAminoAcidDBSize (DBReader.cpp:541-550): data bytes - 2 * entries
    uint64_t aminoAcidDBSize() const { return dataSize_ - 2 * (uint64_t)entries_.size(); }
    uint64_t dataSize() const { return dataSize_; }

private:
    void inflate(const std::string &name);
    std::vector<DBEntry> entries_;
    std::vector<uint32_t> linear_, linearRank_;
    RawBuf data_;
    uint64_t dataSize_ = 0;
    int dbtype_ = 0;
};

class DBWriter {
public:
    // entries are buffered per writer thread and laid out thread after thread on close(), like the reference's
    // per-thread data files merged by DBWriter::mergeResults; the index is written sorted by key, <name>.dbtype last
    // (the workflow scripts test for it: blastp.sh:31,39)
    void open(const std::string &name, int dbtype, int threads);
    // --compressed: mode of every DBWriter opened afterwards (the flag of the module that is running)
    static void setCompressedDefault(bool on);
    void write(int thread, uint32_t key, const char *payload, size_t len);   // payload without the NUL
    // writes what parts [first, first + count) hold to the data file now and empties their buffers; no write() to those parts
    // may run concurrently.  Lets a long search put its output on disk batch by batch, overlapped with the GPU work.
    void flush(int first, int count);
    void close();
    ~DBWriter();

private:
    struct Part {
        std::vector<char> data;        // entries not yet in the file
        std::vector<DBEntry> idx;
        size_t flushed = 0;            // idx[0 .. flushed) carry file offsets, the rest offsets into data
    };
    std::mutex mu_;
    int fd_ = -1;
    uint64_t fileEnd_ = 0;
    std::string name_;
    int dbtype_ = 0;
    bool compressed_ = false;
    std::vector<Part> parts_;
};

}  // namespace spbhost
