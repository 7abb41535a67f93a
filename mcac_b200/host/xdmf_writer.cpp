// mcac_b200 host layer — HDF5 (heavy data) + XMF (light data) writer of the reference's output layout; see xdmf_writer.hpp.
//
// HDF5 file format, as far as it is used here ("HDF5 File Format Specification Version 1.1"):
//   superblock v0 (96 bytes at offset 0) -> root group symbol table entry (cached B-tree / heap addresses)
//   root group   = object header v1 with one Symbol Table message -> B-tree v1 (group nodes) -> symbol table nodes "SNOD"
//                  (entries sorted by link name) + local heap "HEAP" holding the names
//   dataset      = object header v1 with Dataspace v1 (rank 1), Datatype v1 (IEEE f64 / two's complement i32, i64, little endian),
//                  Fill Value v2 (default) and Data Layout v3 (contiguous: address + size) messages
// Raw data is appended as it arrives; all metadata is written by close().
#include "xdmf_writer.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <sstream>

#include "physical_model.hpp"

namespace mcac {
namespace {
constexpr uint64_t kUndef = ~0ULL;
struct Bytes {
    std::vector<unsigned char> b;
    void u8(unsigned v) { b.push_back((unsigned char)v); }
    void u16(unsigned v) { u8(v & 0xff); u8((v >> 8) & 0xff); }
    void u32(uint32_t v) { for (int i = 0; i < 4; i++) u8((v >> (8 * i)) & 0xff); }
    void u64(uint64_t v) { for (int i = 0; i < 8; i++) u8((unsigned)((v >> (8 * i)) & 0xff)); }
    void str(const char *s, size_t n) { b.insert(b.end(), s, s + n); }
    void zeros(size_t n) { b.insert(b.end(), n, 0); }
    void pad8() { while (b.size() % 8) u8(0); }
    size_t size() const { return b.size(); }
};
// one header message: type, size of the (8-byte padded) data, flags, 3 reserved bytes, data
void message(Bytes &out, unsigned type, const Bytes &data) {
    Bytes d = data;
    d.pad8();
    out.u16(type);
    out.u16((unsigned)d.size());
    out.u8(0);
    out.zeros(3);
    out.b.insert(out.b.end(), d.b.begin(), d.b.end());
}
size_t type_size(H5File::Type t) { return t == H5File::I32 ? 4 : 8; }
Bytes dataset_header(H5File::Type type, uint64_t count, uint64_t address) {
    Bytes msgs;
    {   // Dataspace message, version 1: rank 1, no maximum dimensions
        Bytes m;
        m.u8(1); m.u8(1); m.u8(0); m.zeros(5);
        m.u64(count);
        message(msgs, 0x0001, m);
    }
    {   // Datatype message, version 1
        Bytes m;
        if (type == H5File::F64) {
            m.u8(0x11);                     // version 1, class 1 (floating point)
            m.u8(0x20); m.u8(0x3f); m.u8(0);  // little endian, implied mantissa msb, sign bit 63
            m.u32(8);
            m.u16(0); m.u16(64);            // bit offset, precision
            m.u8(52); m.u8(11); m.u8(0); m.u8(52);  // exponent location / size, mantissa location / size
            m.u32(1023);                    // exponent bias
        } else {
            m.u8(0x10);                     // version 1, class 0 (fixed point)
            m.u8(0x08); m.u8(0); m.u8(0);   // little endian, signed
            m.u32((uint32_t)type_size(type));
            m.u16(0); m.u16((unsigned)(8 * type_size(type)));
        }
        message(msgs, 0x0003, m);
    }
    {   // Fill Value message, version 2: late allocation, written if set, default fill value (defined, size 0)
        Bytes m;
        m.u8(2); m.u8(2); m.u8(2); m.u8(1);
        m.u32(0);
        message(msgs, 0x0005, m);
    }
    {   // Data Layout message, version 3, contiguous
        Bytes m;
        m.u8(3); m.u8(1);
        m.u64(count ? address : kUndef);
        m.u64(count * type_size(type));
        message(msgs, 0x0008, m);
    }
    Bytes h;
    h.u8(1); h.u8(0); h.u16(4); h.u32(1); h.u32((uint32_t)msgs.size()); h.zeros(4);
    h.b.insert(h.b.end(), msgs.b.begin(), msgs.b.end());
    return h;
}
}  // namespace

H5File::H5File(const std::string &p) : path(p) {
    f = std::fopen(path.c_str(), "wb");
    if (!f) throw IOError("Error creating file " + path);
    const std::vector<unsigned char> zero(96, 0);  // room for the superblock
    put(zero.data(), zero.size());
}
H5File::~H5File() {
    try { close(); } catch (...) {}
}
void H5File::put(const void *p, size_t n) {
    if (n && std::fwrite(p, 1, n, f) != n) throw IOError("Error writing " + path);
    pos += n;
}
void H5File::pad8() {
    static const unsigned char z[8] = {0};
    if (pos % 8) put(z, 8 - pos % 8);
}
std::string H5File::add(Type type, const void *data, uint64_t count) {
    if (!f) throw IOError("File is already closed: " + path);
    pad8();
    Item it{"Data" + std::to_string(items.size()), type, count, pos};
    put(data, (size_t)(count * type_size(type)));
    items.push_back(it);
    return it.name;
}
void H5File::close() {
    if (!f) return;
    pad8();
    // ---- dataset object headers
    std::vector<uint64_t> header_addr(items.size());
    for (size_t i = 0; i < items.size(); i++) {
        const Bytes h = dataset_header(items[i].type, items[i].count, items[i].address);
        header_addr[i] = pos;
        put(h.b.data(), h.size());
    }
    // ---- local heap: "" at offset 0, then the names (8-byte aligned)
    std::vector<size_t> order(items.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return items[a].name < items[b].name; });  // strcmp order
    Bytes heap_data;
    heap_data.zeros(8);
    std::vector<uint64_t> name_off(items.size());
    for (size_t i : order) {
        name_off[i] = heap_data.size();
        heap_data.str(items[i].name.c_str(), items[i].name.size() + 1);
        heap_data.pad8();
    }
    const uint64_t heap_addr = pos;
    {
        Bytes h;
        h.str("HEAP", 4); h.u8(0); h.zeros(3);
        h.u64(heap_data.size());
        h.u64(1);  // H5HL_FREE_NULL: no free block
        h.u64(heap_addr + 32);
        put(h.b.data(), h.size());
        put(heap_data.b.data(), heap_data.size());
    }
    // ---- symbol table nodes (2 * leaf_k entries each) and one B-tree node over them
    const size_t n = items.size();
    size_t leaf_k = std::max<size_t>(4, (n + 1) / 2);
    leaf_k = std::min<size_t>(leaf_k, 32767);
    const size_t per = 2 * leaf_k, n_snod = std::max<size_t>(1, (n + per - 1) / per);
    const size_t internal_k = std::max<size_t>(16, (n_snod + 1) / 2);
    if (internal_k > 32767) throw IOError("too many datasets for one group: " + path);
    std::vector<uint64_t> snod_addr(n_snod), last_name(n_snod, 0);
    for (size_t s = 0; s < n_snod; s++) {
        const size_t lo = s * per, hi = std::min(n, lo + per);
        Bytes b;
        b.str("SNOD", 4); b.u8(1); b.u8(0); b.u16((unsigned)(hi - lo));
        for (size_t e = lo; e < lo + per; e++) {
            if (e < hi) {
                const size_t i = order[e];
                b.u64(name_off[i]); b.u64(header_addr[i]); b.u32(0); b.u32(0); b.zeros(16);
                last_name[s] = name_off[i];
            } else b.zeros(40);
        }
        snod_addr[s] = pos;
        put(b.b.data(), b.size());
    }
    const uint64_t btree_addr = pos;
    {
        Bytes b;
        b.str("TREE", 4); b.u8(0); b.u8(0); b.u16((unsigned)(n ? n_snod : 0));
        b.u64(kUndef); b.u64(kUndef);
        b.u64(0);  // key 0: the empty name
        for (size_t s = 0; s < 2 * internal_k; s++) {
            if (s < n_snod && n) { b.u64(snod_addr[s]); b.u64(last_name[s]); }
            else { b.u64(0); b.u64(0); }
        }
        put(b.b.data(), b.size());
    }
    // ---- root group object header: one Symbol Table message
    const uint64_t root_addr = pos;
    {
        Bytes msgs, m;
        m.u64(btree_addr); m.u64(heap_addr);
        message(msgs, 0x0011, m);
        Bytes h;
        h.u8(1); h.u8(0); h.u16(1); h.u32(1); h.u32((uint32_t)msgs.size()); h.zeros(4);
        h.b.insert(h.b.end(), msgs.b.begin(), msgs.b.end());
        put(h.b.data(), h.size());
    }
    const uint64_t eof = pos;
    // ---- superblock, version 0
    Bytes sb;
    static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    sb.str(reinterpret_cast<const char *>(sig), 8);
    sb.u8(0); sb.u8(0); sb.u8(0); sb.u8(0); sb.u8(0);  // versions: superblock, free space, root symbol table; reserved; shared header
    sb.u8(8); sb.u8(8); sb.u8(0);                      // size of offsets, size of lengths, reserved
    sb.u16((unsigned)leaf_k); sb.u16((unsigned)internal_k);
    sb.u32(0);                                         // file consistency flags
    sb.u64(0); sb.u64(kUndef); sb.u64(eof); sb.u64(kUndef);  // base, free-space info, end of file, driver info
    sb.u64(0); sb.u64(root_addr); sb.u32(1); sb.u32(0); sb.u64(btree_addr); sb.u64(heap_addr);  // root symbol table entry
    std::fseek(f, 0, SEEK_SET);
    const bool ok = std::fwrite(sb.b.data(), 1, sb.size(), f) == sb.size();
    const bool closed = std::fclose(f) == 0;
    f = nullptr;
    if (!ok || !closed) throw IOError("Error writing " + path);
}

// ------------------------------------------------------------------------------------------------------------------------------
std::string XdmfSeriesWriter::filename(int step, size_t n) {
    const int width = int(std::ceil(std::log10(static_cast<float>(n)))) + 4;
    std::ostringstream s;
    s << "_" << std::setfill('0') << std::setw(width) << step;
    return s.str();
}
XdmfSeriesWriter::XdmfSeriesWriter(std::string p, std::string g, size_t per_file, size_t n_width, std::vector<std::pair<std::string, std::string>> ph)
    : prefix(std::move(p)), grid_name(std::move(g)), n_time_per_file(std::max<size_t>(1, per_file)), n_for_width(n_width), physics(std::move(ph)) {}
XdmfSeriesWriter::~XdmfSeriesWriter() {
    try { flush(); } catch (...) {}
}
void XdmfSeriesWriter::open_file() {
    const std::string base = prefix + filename(num_file, n_for_width);
    h5 = new H5File(base + ".h5");
    const size_t slash = base.find_last_of('/');
    h5_basename = (slash == std::string::npos ? base : base.substr(slash + 1)) + ".h5";
    xml_grids.clear();
}
std::string XdmfSeriesWriter::data_item(H5File::Type type, uint64_t count, const std::string &dataset) const {
    std::ostringstream s;
    s << "<DataItem DataType=\"" << (type == H5File::F64 ? "Float" : "Int") << "\" Dimensions=\"" << count << "\" Format=\"HDF\" Precision=\""
      << (type == H5File::I32 ? 4 : 8) << "\">" << h5_basename << ":" << dataset << "</DataItem>";
    return s.str();
}
void XdmfSeriesWriter::begin_step(double time) {
    if (step % n_time_per_file == 0 && !h5) open_file();
    std::ostringstream s;
    s << std::setprecision(17) << time;
    cur = "      <Grid Name=\"" + grid_name + "\">\n        <Time Value=\"" + s.str() + "\"/>\n";
}
void XdmfSeriesWriter::positions(const double *xyz, uint64_t n_points) {
    const std::string ds = h5->add(H5File::F64, xyz, 3 * n_points);
    cur += "        <Geometry Origin=\"\" Type=\"XYZ\">\n          " + data_item(H5File::F64, 3 * n_points, ds) + "\n        </Geometry>\n";
    cur += "        <Topology Dimensions=\"" + std::to_string(n_points) + "\" Type=\"Polyvertex\"/>\n";
}
void XdmfSeriesWriter::attribute(const std::string &name, H5File::Type type, const void *data, uint64_t count, bool scalar_on_nodes) {
    const std::string ds = h5->add(type, data, count);
    cur += "        <Attribute Center=\"Node\" ElementCell=\"\" ElementDegree=\"0\" ElementFamily=\"\" ItemType=\"\" Name=\"" + name + "\" Type=\"" +
           (scalar_on_nodes ? "Scalar" : "None") + "\">\n          " + data_item(type, count, ds) + "\n        </Attribute>\n";
}
void XdmfSeriesWriter::end_step() {
    xml_grids += cur + "      </Grid>\n";
    cur.clear();
    step++;
    if (step % n_time_per_file == 0) flush();
}
void XdmfSeriesWriter::flush() {
    if (!h5) return;
    H5File *file = h5;
    h5 = nullptr;
    file->close();
    delete file;
    const std::string base = prefix + filename(num_file, n_for_width);
    std::FILE *x = std::fopen((base + ".xmf").c_str(), "w");
    if (!x) throw IOError("Error creating file " + base + ".xmf");
    std::string s = "<?xml version=\"1.0\" encoding=\"utf-8\"?>\n<!DOCTYPE Xdmf SYSTEM \"Xdmf.dtd\" []>\n"
                    "<Xdmf xmlns:xi=\"http://www.w3.org/2001/XInclude\" Version=\"3.0\">\n  <Domain>\n"
                    "    <Information Name=\"Copyright\" Value=\"Produced by MCAC\"/>\n"
                    "    <Information Name=\"Physics\" Value=\"Physical properties of the simulation\">\n";
    for (const auto &kv : physics) s += "      <Information Name=\"" + kv.first + "\" Value=\"" + kv.second + "\"/>\n";
    s += "    </Information>\n    <Grid CollectionType=\"Temporal\" GridType=\"Collection\" Name=\"Collection\">\n" + xml_grids + "    </Grid>\n  </Domain>\n</Xdmf>\n";
    const bool ok = std::fwrite(s.data(), 1, s.size(), x) == s.size();
    if (std::fclose(x) != 0 || !ok) throw IOError("Error writing " + base + ".xmf");
    xml_grids.clear();
    num_file++;
}

}  // namespace mcac
