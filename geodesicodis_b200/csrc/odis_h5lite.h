// Minimal HDF5 writer for DATA/data.h5.
//
// The reference creates every dataset with H5P_DEFAULT property lists, H5T_NATIVE_FLOAT, a fixed shape
// and no attributes (/root/reference/src/outFiles.cpp:204-339,499-515) and writes one row per dump through
// a hyperslab (:522-684). That needs only the oldest on-disk dialect, which is written here directly
// (no HDF5 library exists in this image): superblock version 0, a root group kept as a symbol table
// (one v1 B-tree node, one local heap, one symbol node), and per dataset a version-1 object header with
// dataspace, datatype (IEEE f32 LE), fill-value, contiguous-layout and modification-time messages — the
// same dialect the HDF5 library itself produced for the reference's shipped input_files/*.h5 files.
// All datasets are declared first; the data region of each is allocated up front (sparse file), so a
// dump is one pwrite per dataset.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace odis {

class H5LiteWriter {
public:
    H5LiteWriter() = default;
    ~H5LiteWriter();
    H5LiteWriter(const H5LiteWriter&) = delete;
    H5LiteWriter& operator=(const H5LiteWriter&) = delete;

    // Truncates/creates the file (H5F_ACC_TRUNC). Returns 0 or -1 (err set).
    int create(const std::string& path, std::string& err);
    // Declares a float32 dataset of rank 1 or 2; returns its index, or -1 (duplicate name, bad rank,
    // or already finalised). At most 32 datasets.
    int add_dataset(const std::string& name, int rank, const uint64_t* dims, std::string& err);
    // Writes all metadata and sizes the file. Called implicitly by the first write.
    int finalize(std::string& err);
    // Writes `count` floats starting at row `row` (rank 2: row-major rows of dims[1]; rank 1: element index).
    int write_rows(int dataset, uint64_t row, uint64_t nrows, const float* data, std::string& err);
    int close(std::string& err);
    bool is_open() const { return fd_ >= 0; }

private:
    struct Dataset {
        std::string name;
        int rank = 0;
        uint64_t dims[2] = {0, 0};
        uint64_t header_addr = 0, data_addr = 0, data_bytes = 0, heap_offset = 0;
    };
    int fd_ = -1;
    bool finalized_ = false;
    std::vector<Dataset> sets_;
};

}  // namespace odis
