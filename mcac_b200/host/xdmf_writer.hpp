// mcac_b200 host layer — the reference's output files (src/io/*: XDMF light data + HDF5 heavy data), written without libXdmf /
// libhdf5 (neither exists in this image): a minimal HDF5 writer (superblock version 0, one root group, contiguous 1-D datasets, no
// filters — exactly what XdmfHDF5Writer produces with deflate off, writer.cpp:125-135) and the XMF tree of
// io/aggregat_list.cpp:36-67, io/sphere_list.cpp:36-57, io/physical_model.cpp:30-49 as pymcac/reader/xdmf_reader.py:58-112 and
// h5_reader.py:56-177 read it.  File naming and rotation follow ThreadedIO (threaded_io.cpp, writer.cpp:72-112, format.cpp:60-65):
// <prefix>_<k>.{h5,xmf}, k zero-padded to ceil(log10(N)) + 4 digits, n_time_per_file grids per file.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

namespace mcac {

// One HDF5 file holding datasets "Data0", "Data1", ... in its root group (the names XdmfHDF5Writer gives its heavy data).
class H5File {
  public:
    enum Type { F64, I32, I64 };
    explicit H5File(const std::string &path);  // throws IOError
    ~H5File();
    H5File(const H5File &) = delete;
    H5File &operator=(const H5File &) = delete;
    // appends the raw data now, describes it at close(); returns the dataset's name
    std::string add(Type type, const void *data, uint64_t count);
    void close();  // writes object headers, local heap, symbol table node, B-tree, root header and the superblock

  private:
    struct Item { std::string name; Type type; uint64_t count, address; };
    std::string path;
    std::FILE *f = nullptr;
    uint64_t pos = 0;
    std::vector<Item> items;
    void put(const void *p, size_t n);
    void pad8();
};

// The writer of one series (the reference's ThreadedIO for "Spheres" or "Aggregats")
class XdmfSeriesWriter {
  public:
    // physics: the 16 name / value pairs of PhysicalModel::xmf_write in order; n_for_width: N of filename(step, N)
    XdmfSeriesWriter(std::string prefix, std::string grid_name, size_t n_time_per_file, size_t n_for_width,
                     std::vector<std::pair<std::string, std::string>> physics);
    ~XdmfSeriesWriter();  // flushes an unfinished file, like ~ThreadedIO
    void begin_step(double time);
    void positions(const double *xyz_interleaved, uint64_t n_points);
    void attribute(const std::string &name, H5File::Type type, const void *data, uint64_t count, bool scalar_on_nodes = true);
    void end_step();
    void flush();
    static std::string filename(int step, size_t n);  // format.cpp:60-65

  private:
    std::string prefix, grid_name;
    size_t n_time_per_file, n_for_width;
    std::vector<std::pair<std::string, std::string>> physics;
    size_t step = 0;
    int num_file = 0;
    H5File *h5 = nullptr;
    std::string h5_basename, xml_grids, cur;
    void open_file();
    std::string data_item(H5File::Type type, uint64_t count, const std::string &dataset) const;
};

}  // namespace mcac
