// HDF5 superblock-v0 writer (see odis_h5lite.h). Field layouts follow the HDF5 File Format
// Specification, version 1.1 structures (the ones a default H5Fcreate/H5Dcreate produces).
#include "odis_h5lite.h"

#include <algorithm>
#include <cstring>
#include <ctime>
#include <fcntl.h>
#include <unistd.h>

namespace odis {
namespace {

constexpr uint64_t kUndef = ~0ull;
constexpr int kLeafK = 16;        // symbol-table node holds up to 2*kLeafK = 32 entries (H5Pset_sym_k analogue)
constexpr int kInternalK = 16;
constexpr uint64_t kDataAlign = 4096;

struct Buf {
    std::vector<uint8_t> b;
    void u8(uint8_t v) { b.push_back(v); }
    void u16(uint16_t v) { for (int i = 0; i < 2; i++) b.push_back((uint8_t)(v >> (8 * i))); }
    void u32(uint32_t v) { for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i))); }
    void u64(uint64_t v) { for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i))); }
    void bytes(const void* p, size_t n) { const uint8_t* q = (const uint8_t*)p; b.insert(b.end(), q, q + n); }
    void zeros(size_t n) { b.insert(b.end(), n, 0); }
    void pad8() { while (b.size() % 8) b.push_back(0); }
    size_t size() const { return b.size(); }
};

uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// version-1 object header of one float32 dataset
Buf dataset_header(int rank, const uint64_t* dims, uint64_t data_addr, uint64_t data_bytes, uint32_t mtime) {
    Buf m;   // messages
    auto msg = [&](uint16_t type, const Buf& body) {
        Buf padded = body;
        padded.pad8();
        m.u16(type); m.u16((uint16_t)padded.size()); m.u8(0); m.zeros(3);
        m.bytes(padded.b.data(), padded.size());
    };
    {   // 0x0001 dataspace, version 1, max dims present
        Buf d; d.u8(1); d.u8((uint8_t)rank); d.u8(1); d.zeros(5);
        for (int i = 0; i < rank; i++) d.u64(dims[i]);
        for (int i = 0; i < rank; i++) d.u64(dims[i]);
        msg(0x0001, d);
    }
    {   // 0x0003 datatype: class 1 (floating point) version 1, IEEE binary32 little endian
        Buf d; d.u8(0x11); d.u8(0x20); d.u8(0x1f); d.u8(0x00); d.u32(4);
        d.u16(0); d.u16(32); d.u8(23); d.u8(8); d.u8(0); d.u8(23); d.u32(127);
        msg(0x0003, d);
    }
    {   // 0x0005 fill value, version 2: allocate late / write if set / undefined-default, size 0
        Buf d; d.u8(2); d.u8(2); d.u8(2); d.u8(1); d.u32(0);
        msg(0x0005, d);
    }
    {   // 0x0008 data layout, version 3, class 1 contiguous
        Buf d; d.u8(3); d.u8(1); d.u64(data_addr); d.u64(data_bytes);
        msg(0x0008, d);
    }
    {   // 0x0012 object modification time, version 1
        Buf d; d.u8(1); d.zeros(3); d.u32(mtime);
        msg(0x0012, d);
    }
    Buf h;
    h.u8(1); h.u8(0); h.u16(5); h.u32(1); h.u32((uint32_t)m.size());
    h.zeros(4);                    // prefix is padded to 16 bytes
    h.bytes(m.b.data(), m.size());
    return h;
}

}  // namespace

H5LiteWriter::~H5LiteWriter() {
    std::string e;
    close(e);
}

int H5LiteWriter::create(const std::string& path, std::string& err) {
    if (fd_ >= 0) { err = "file already open"; return -1; }
    fd_ = ::open(path.c_str(), O_CREAT | O_TRUNC | O_RDWR, 0644);
    if (fd_ < 0) { err = "cannot create " + path; return -1; }
    finalized_ = false;
    sets_.clear();
    return 0;
}

int H5LiteWriter::add_dataset(const std::string& name, int rank, const uint64_t* dims, std::string& err) {
    if (fd_ < 0) { err = "file not open"; return -1; }
    if (finalized_) { err = "datasets must be declared before the first write"; return -1; }
    if (rank < 1 || rank > 2) { err = "rank must be 1 or 2"; return -1; }
    if (name.empty()) { err = "empty dataset name"; return -1; }
    if ((int)sets_.size() >= 2 * kLeafK) { err = "too many datasets"; return -1; }
    for (const auto& d : sets_)
        if (d.name == name) { err = "dataset '" + name + "' already exists"; return -1; }   // HDF5 refuses duplicates too
    Dataset d;
    d.name = name; d.rank = rank;
    d.dims[0] = dims[0]; d.dims[1] = rank == 2 ? dims[1] : 1;
    d.data_bytes = d.dims[0] * d.dims[1] * 4;
    sets_.push_back(d);
    return (int)sets_.size() - 1;
}

int H5LiteWriter::finalize(std::string& err) {
    if (fd_ < 0) { err = "file not open"; return -1; }
    if (finalized_) return 0;
    const uint32_t now = (uint32_t)std::time(nullptr);
    // ---- layout ----
    const uint64_t root_header = 96;                                   // superblock is 96 bytes
    const uint64_t btree_addr = root_header + 16 + 24;                 // v1 header prefix + one symbol-table message
    const uint64_t btree_size = 8 + 16 + (2 * kInternalK + 1) * 8 + 2 * kInternalK * 8;
    const uint64_t heap_addr = btree_addr + btree_size;
    uint64_t heap_data = 8;                                            // offset 0: the root's empty name
    for (auto& d : sets_) { d.heap_offset = heap_data; heap_data += align_up(d.name.size() + 1, 8); }
    const uint64_t heap_free_off = heap_data;
    heap_data += 16;                                                   // one trailing free block
    const uint64_t heap_data_addr = heap_addr + 32;
    const uint64_t snod_addr = align_up(heap_data_addr + heap_data, 8);
    const uint64_t snod_size = 8 + 2 * kLeafK * 40;
    uint64_t pos = snod_addr + snod_size;
    std::vector<Buf> headers(sets_.size());
    for (size_t i = 0; i < sets_.size(); i++) {
        sets_[i].header_addr = pos;
        pos += dataset_header(sets_[i].rank, sets_[i].dims, 0, 0, now).size();
    }
    for (auto& d : sets_) {
        pos = align_up(pos, kDataAlign);
        d.data_addr = pos;
        pos += d.data_bytes;
    }
    const uint64_t eof = pos;
    for (size_t i = 0; i < sets_.size(); i++) headers[i] = dataset_header(sets_[i].rank, sets_[i].dims, sets_[i].data_addr, sets_[i].data_bytes, now);

    Buf f;
    // ---- superblock, version 0 ----
    const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    f.bytes(sig, 8);
    f.u8(0); f.u8(0); f.u8(0); f.u8(0); f.u8(0); f.u8(8); f.u8(8); f.u8(0);
    f.u16(kLeafK); f.u16(kInternalK); f.u32(0);
    f.u64(0); f.u64(kUndef); f.u64(eof); f.u64(kUndef);
    // root group symbol table entry
    f.u64(0); f.u64(root_header); f.u32(1); f.u32(0); f.u64(btree_addr); f.u64(heap_addr);
    // ---- root object header (v1): one symbol-table message ----
    f.u8(1); f.u8(0); f.u16(1); f.u32(1); f.u32(24); f.zeros(4);
    f.u16(0x0011); f.u16(16); f.u8(0); f.zeros(3); f.u64(btree_addr); f.u64(heap_addr);
    // ---- group B-tree node (leaf, one child) ----
    std::vector<size_t> order(sets_.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return sets_[a].name < sets_[b].name; });
    f.bytes("TREE", 4); f.u8(0); f.u8(0); f.u16(sets_.empty() ? 0 : 1); f.u64(kUndef); f.u64(kUndef);
    {
        const size_t start = f.size();
        if (!sets_.empty()) {
            f.u64(0);                                   // key 0: the empty string
            f.u64(snod_addr);
            f.u64(sets_[order.back()].heap_offset);     // key 1: largest name in the child
        }
        f.zeros(((2 * kInternalK + 1) + 2 * kInternalK) * 8 - (f.size() - start));
    }
    // ---- local heap ----
    f.bytes("HEAP", 4); f.u8(0); f.zeros(3); f.u64(heap_data); f.u64(heap_free_off); f.u64(heap_data_addr);
    f.zeros(8);
    for (const auto& d : sets_) {
        f.bytes(d.name.c_str(), d.name.size() + 1);
        f.pad8();
    }
    f.u64(1); f.u64(16);                                // free block: no next (1 = H5HL_FREE_NULL), 16 bytes
    f.pad8();
    // ---- symbol node ----
    f.bytes("SNOD", 4); f.u8(1); f.u8(0); f.u16((uint16_t)sets_.size());
    for (size_t k : order) { f.u64(sets_[k].heap_offset); f.u64(sets_[k].header_addr); f.u32(0); f.u32(0); f.zeros(16); }
    f.zeros((2 * kLeafK - sets_.size()) * 40);
    for (const auto& h : headers) f.bytes(h.b.data(), h.size());

    if (::pwrite(fd_, f.b.data(), f.size(), 0) != (ssize_t)f.size() || ::ftruncate(fd_, (off_t)eof) != 0) {
        err = "write failed";
        return -1;
    }
    finalized_ = true;
    return 0;
}

int H5LiteWriter::write_rows(int dataset, uint64_t row, uint64_t nrows, const float* data, std::string& err) {
    if (fd_ < 0) { err = "file not open"; return -1; }
    if (dataset < 0 || dataset >= (int)sets_.size()) { err = "bad dataset index"; return -1; }
    if (!finalized_ && finalize(err) != 0) return -1;
    const Dataset& d = sets_[(size_t)dataset];
    const uint64_t row_elems = d.rank == 2 ? d.dims[1] : 1;
    if (row + nrows > d.dims[0]) { err = "selection outside the dataset extent"; return -1; }   // HDF5: H5Sselect_hyperslab fails
    const size_t bytes = (size_t)(nrows * row_elems * 4);
    if (::pwrite(fd_, data, bytes, (off_t)(d.data_addr + row * row_elems * 4)) != (ssize_t)bytes) { err = "write failed"; return -1; }
    return 0;
}

int H5LiteWriter::close(std::string& err) {
    if (fd_ < 0) return 0;
    int rc = 0;
    if (!finalized_) rc = finalize(err);
    ::close(fd_);
    fd_ = -1;
    return rc;
}

}  // namespace odis
