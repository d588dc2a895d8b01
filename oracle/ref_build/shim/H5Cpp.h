// TEST INFRASTRUCTURE — not product code.
//
// From-scratch stand-in for the handful of HDF5 C entry points the GeodesicODIS
// reference calls (H5Fcreate / H5Screate_simple / H5Dcreate / H5Sselect_hyperslab /
// H5Dwrite; /root/reference/src/outFiles.cpp:204-339,499-515,557-677). The image has
// no HDF5 library. Datasets are held in memory as float32 and written, one raw
// little-endian file per dataset plus an index, by h5shim::dump_all(dir) so the
// test-suite can compare them against the product's own data.h5.
#ifndef ODIS_ORACLE_H5_SHIM_H
#define ODIS_ORACLE_H5_SHIM_H

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

typedef int64_t hid_t;
typedef unsigned long long hsize_t;
typedef int herr_t;

#define H5F_ACC_TRUNC 2u
#define H5P_DEFAULT ((hid_t)0)
#define H5T_NATIVE_FLOAT ((hid_t)-101)
#define H5T_NATIVE_DOUBLE ((hid_t)-102)
#define H5T_NATIVE_INT ((hid_t)-103)
enum H5S_seloper_t { H5S_SELECT_SET = 0 };

namespace H5 {}

namespace h5shim {
struct Space { std::vector<hsize_t> dims, start, count; bool selected = false; };
struct Dataset { std::string name; std::vector<hsize_t> dims; std::vector<float> data; };
struct Registry {
    std::vector<Space> spaces;
    std::vector<Dataset> sets;
    std::string file;
};
inline Registry& reg() { static Registry r; return r; }

inline void dump_all(const char* dir) {
    Registry& r = reg();
    std::string idx = std::string(dir) + "/h5shim_index.txt";
    FILE* fi = std::fopen(idx.c_str(), "w");
    if (!fi) return;
    for (size_t k = 0; k < r.sets.size(); k++) {
        const Dataset& d = r.sets[k];
        std::string fn = std::string(dir) + "/h5shim_" + std::to_string(k) + ".f32";
        FILE* f = std::fopen(fn.c_str(), "wb");
        if (f) { std::fwrite(d.data.data(), sizeof(float), d.data.size(), f); std::fclose(f); }
        std::fprintf(fi, "%zu|%s|", k, d.name.c_str());
        for (size_t j = 0; j < d.dims.size(); j++) std::fprintf(fi, "%s%llu", j ? "," : "", d.dims[j]);
        std::fprintf(fi, "\n");
    }
    std::fclose(fi);
}
}  // namespace h5shim

inline hid_t H5Fcreate(const char* name, unsigned, hid_t, hid_t) {
    h5shim::reg().file = name;
    return 1;
}
inline hid_t H5Screate_simple(int rank, const hsize_t* dims, const hsize_t*) {
    h5shim::Space s; s.dims.assign(dims, dims + rank);
    h5shim::reg().spaces.push_back(s);
    return (hid_t)h5shim::reg().spaces.size() - 1;
}
inline hid_t H5Dcreate(hid_t, const char* name, hid_t, hid_t space, hid_t, hid_t, hid_t) {
    h5shim::Dataset d; d.name = name; d.dims = h5shim::reg().spaces[space].dims;
    size_t n = 1; for (hsize_t x : d.dims) n *= (size_t)x;
    d.data.assign(n, 0.0f);
    h5shim::reg().sets.push_back(d);
    return (hid_t)h5shim::reg().sets.size() - 1;
}
inline herr_t H5Sselect_hyperslab(hid_t space, H5S_seloper_t, const hsize_t* start, const hsize_t*,
                                  const hsize_t* count, const hsize_t*) {
    h5shim::Space& s = h5shim::reg().spaces[space];
    s.start.assign(start, start + s.dims.size());
    s.count.assign(count, count + s.dims.size());
    s.selected = true;
    return 0;
}
inline herr_t H5Dwrite(hid_t set, hid_t, hid_t, hid_t filespace, hid_t, const void* buf) {
    h5shim::Dataset& d = h5shim::reg().sets[set];
    const h5shim::Space& s = h5shim::reg().spaces[filespace];
    const float* src = (const float*)buf;
    const size_t rank = d.dims.size();
    if (!s.selected) { std::memcpy(d.data.data(), src, d.data.size() * sizeof(float)); return 0; }
    // a selection outside the dataset extent is an error in HDF5 (nothing is written)
    for (size_t k = 0; k < rank; k++)
        if (s.start[k] + s.count[k] > d.dims[k]) return -1;
    if (rank == 1) {
        for (hsize_t i = 0; i < s.count[0]; i++) d.data[s.start[0] + i] = src[i];
    } else if (rank == 2) {
        size_t k = 0;
        for (hsize_t i = 0; i < s.count[0]; i++)
            for (hsize_t j = 0; j < s.count[1]; j++)
                d.data[(s.start[0] + i) * d.dims[1] + s.start[1] + j] = src[k++];
    }
    return 0;
}
#endif
