// Output<LT, T>: VTK output of the fields a main holds after a write interval -- the consumer of the
// path's results.  Same public interface and, byte for byte, the same files as the reference's
// Output / VTK::Output (src/io/Output.h:26-174, src/io/VTK.h): per rank an UnstructuredGrid .vtu
// (<dir>/vtu/RRRR_<name>_NNNNNNN.vtu: voxel / pixel cells centred on the node positions, raw appended
// binary, every array preceded by its UInt32 byte count) and one <dir>/<name>_NNNNNNN.pvtu that lists
// the pieces in fixed 100-byte records.
//
// Differences in construction, not in output: the mesh is built once on a dense corner lattice of
// the bounding box; every array is a (source pointer, element type, index list) triple converted while
// it is written; the .pvtu needs neither MPI-IO nor a broadcast -- every rank formats the same header,
// so it knows where its piece record goes and writes it with pwrite (this also avoids the 101-byte
// sprintf overflow of VTK.h:1270-1274 in the reference's non-MPI branch).
#ifndef CHIMP_OUTPUT_H
#define CHIMP_OUTPUT_H

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>
#include <ctime>
#include <functional>
#include <iomanip>
#include <sstream>

#include "LBfield.h"
#include "LBgrid.h"

namespace VTK {
static constexpr int BINARY = 0;
static constexpr int ASCII = 1;
struct voxel {};
} // namespace VTK

namespace chimp_host {

template <typename T> struct VtkTypeName;
template <> struct VtkTypeName<int> { static const char *name() { return "Int32"; } };
template <> struct VtkTypeName<unsigned int> { static const char *name() { return "UInt32"; } };
template <> struct VtkTypeName<float> { static const char *name() { return "Float32"; } };
template <> struct VtkTypeName<double> { static const char *name() { return "Float64"; } };

// Cells of the output mesh: one voxel (3-D, VTK type 11) or pixel (2-D, type 8) per node, corners at
// node + {0,1}^dim - 1/2 (VTK.h:200-256).  Corner points are numbered in order of first appearance
// while the cells are visited in node order (VTK.h:494-557), points are always written with 3 components.
struct CellMesh {
    std::vector<float> points;
    std::vector<int> connectivity, offsets, types;
    int nCells = 0, nPoints = 0;

    CellMesh(const std::vector<int> &nodePos, int dim)
    {
        const int nCorners = 1 << dim;
        nCells = int(nodePos.size()) / dim;
        std::vector<int> lo(dim, 999999), hi(dim, -1);
        for (int n = 0; n < nCells; ++n)
            for (int d = 0; d < dim; ++d) {
                lo[d] = std::min(lo[d], nodePos[n * dim + d]);
                hi[d] = std::max(hi[d], nodePos[n * dim + d]);
            }
        std::vector<long long> stride(dim, 1);
        for (int d = 1; d < dim; ++d) stride[d] = stride[d - 1] * (hi[d - 1] - lo[d - 1] + 2);
        const long long nLattice = nCells ? stride[dim - 1] * (hi[dim - 1] - lo[dim - 1] + 2) : 0;
        std::vector<int> pointOf(std::size_t(nLattice), -1);
        connectivity.reserve(std::size_t(nCells) * nCorners);
        for (int n = 0; n < nCells; ++n)
            for (int corner = 0; corner < nCorners; ++corner) {
                long long at = 0;
                int coord[3] = {0, 0, 0};
                for (int d = 0; d < dim; ++d) {
                    coord[d] = nodePos[n * dim + d] + ((corner >> d) & 1); // x fastest: {0,0,0},{1,0,0},{0,1,0},...
                    at += (coord[d] - lo[d]) * stride[d];
                }
                if (pointOf[at] < 0) {
                    pointOf[at] = nPoints++;
                    for (int d = 0; d < 3; ++d) points.push_back(d < dim ? float(coord[d] - 0.5) : 0.0f);
                }
                connectivity.push_back(pointOf[at]);
            }
        offsets.resize(nCells);
        for (int n = 0; n < nCells; ++n) offsets[n] = nCorners * (n + 1);
        types.assign(nCells, dim == 3 ? 11 : 8);
    }
};

} // namespace chimp_host

template <typename LT, typename T = double, int FMT = VTK::BINARY, typename CELL = VTK::voxel>
class Output
{
public:
    // a variable over a node list of a grid (Output.h:51-53)
    Output(const Grid<LT> &grid, std::vector<int> &nodes, const std::string &dir, int rank, int nproc)
        : mesh_(grid.pos(nodes), LT::nD), dir_(withSlash(dir)), rank_(rank), maxRank_(nproc - 1), nodes_(&nodes)
    {
    }
    // positions given directly (Output.h:46-48), optionally with one std::vector variable (Output.h:36-42)
    Output(const std::vector<int> &pos, const std::string &dir, int rank, int nproc)
        : mesh_(pos, LT::nD), dir_(withSlash(dir)), rank_(rank), maxRank_(nproc - 1)
    {
    }
    Output(const std::vector<int> &pos, const std::string &dir, int rank, int nproc, const std::string &varname, const std::vector<T> &var)
        : Output(pos, dir, rank, nproc)
    {
        add_file(varname);
        addArray(varname, [&var](int i) { return var[i]; }, contiguous(int(var.size())));
    }

    void add_file(const std::string &name) { files_.push_back({name, {}, 0}); }

    // names[i] belongs to fields[i]; multi-field variables get the field number appended (Output.h:79-88)
    template <typename F>
    void add_variables(const std::vector<std::string> &names, const std::vector<std::reference_wrapper<const F>> &fields)
    {
        for (std::size_t i = 0; i < names.size(); ++i)
            for (int f = 0; f < fields[i].get().num_fields(); ++f)
                addField(f, fields[i].get().num_fields() > 1 ? names[i] + std::to_string(f) : names[i], fields[i].get());
    }
    void add_scalar_variables(const std::vector<std::string> &names, const std::vector<std::reference_wrapper<const ScalarField>> &fields)
    {
        add_variables<ScalarField>(names, fields);
    }
    void add_vector_variables(const std::vector<std::string> &names, const std::vector<std::reference_wrapper<const VectorField<LT>>> &fields)
    {
        add_variables<VectorField<LT>>(names, fields);
    }
    template <typename F>
    void add_variable(const std::string &name, const F &field)
    {
        add_variables<F>({name}, {field});
    }
    template <typename F>
    void add_variable_with_names(const std::vector<std::string> &names, const F &field)
    {
        if (int(names.size()) != field.num_fields())
            chimp_host::die("in add_variables: Size of names-vector does not match number of fields, " + std::to_string(names.size()) +
                            " != " + std::to_string(field.num_fields()));
        for (int i = 0; i < field.num_fields(); ++i) addField(i, names[i], field);
    }
    // plain std::vector variables, one value (or nD values) per cell (Output.h:140-146)
    void add_variables(const std::vector<std::string> &names, const std::vector<std::reference_wrapper<std::vector<T>>> &vectors)
    {
        for (std::size_t i = 0; i < names.size(); ++i) {
            const std::vector<T> &v = vectors[i].get();
            addArray(names[i], [&v](int k) { return v[k]; }, contiguous(int(v.size())));
        }
    }

    void write(double t = 0.0)
    {
        for (auto &f : files_) {
            writeVtu(f);
            writePvtu(f, t);
            ++f.nWrite;
        }
    }

private:
    struct Array {
        std::string name;
        int nComp;                   // components per cell as written (2-D vectors are padded to 3, VTK.h:664-673)
        std::function<T(int)> at;    // element i of the source container, converted to the on-disk type
        std::vector<int> index;      // element per written value, -1 writes zero
    };
    struct File {
        std::string name;
        std::vector<Array> arrays;
        int nWrite;
    };

    static std::string withSlash(std::string d)
    {
        if (d.empty() || d.back() != '/') d += '/';
        return d;
    }
    static std::vector<int> contiguous(int n)
    {
        std::vector<int> v(n);
        for (int i = 0; i < n; ++i) v[i] = i;
        return v;
    }
    template <typename F>
    void addField(int fieldNo, const std::string &name, const F &field)
    {
        if (!nodes_) chimp_host::die("Output: field variables need the constructor with a node list");
        std::vector<int> ind;
        ind.reserve(nodes_->size() * field.dim());
        for (int node : *nodes_)
            for (int d = 0; d < field.dim(); ++d) ind.push_back(field.index(fieldNo, d, node));
        const lbBase_t *src = field.data();
        addArray(name, [src](int i) { return T(src[i]); }, ind);
    }
    void addArray(const std::string &name, std::function<T(int)> at, std::vector<int> index)
    {
        if (files_.empty()) chimp_host::die("in VTK::Output::add_variable(" + name + "): Add an output file using 'add_file(name)' before adding variables");
        int nComp = mesh_.nCells ? int(index.size()) / mesh_.nCells : 1;
        if (nComp != 1 && nComp != LT::nD)
            chimp_host::die("in VTK::Output::add_variable(" + name + "): Wrong dimension: Expected " + std::to_string(LT::nD) + " or 1, got " + std::to_string(nComp));
        if (nComp == 2) { // ParaView wants 3-component vectors: a zero third component is written
            std::vector<int> padded(std::size_t(mesh_.nCells) * 3, -1);
            for (int n = 0; n < mesh_.nCells; ++n) { padded[3 * n] = index[2 * n]; padded[3 * n + 1] = index[2 * n + 1]; }
            index.swap(padded);
            nComp = 3;
        }
        files_.back().arrays.push_back({name, nComp, std::move(at), std::move(index)});
    }
    static void makeDir(const std::string &d)
    {
        struct stat st;
        if (stat(d.c_str(), &st) == -1) mkdir(d.c_str(), 0700);
    }
    static std::string arrayTag(const std::string &name, const char *type, int nComp, unsigned offset)
    {
        std::ostringstream s;
        s << "        <DataArray Name=\"" << name << "\" type=\"" << type << "\" NumberOfComponents=\"" << nComp << "\" ";
        if (FMT == VTK::BINARY) s << "format=\"appended\" offset=\"" << offset << "\" ";
        else s << "format=\"ascii\" ";
        s << "> ";
        return s.str();
    }
    template <typename V>
    static void asciiValue(std::ostream &o, V v)
    {
        if (std::abs(double(v)) < 1e-20) o << "0 ";
        else o << v << " ";
    }
    template <typename V>
    static void raw(std::ostream &o, const std::vector<V> &v)
    {
        const unsigned nbytes = unsigned(v.size() * sizeof(V));
        o.write((const char *)&nbytes, sizeof(unsigned));
        o.write((const char *)v.data(), nbytes);
    }
    std::string vtuName(const File &f) const
    {
        std::ostringstream ss;
        ss << std::setfill('0') << std::setw(4) << rank_ << "_" << f.name << "_" << std::setw(7) << f.nWrite << ".vtu";
        return ss.str();
    }

    void writeVtu(const File &f) const
    {
        makeDir(dir_);
        makeDir(dir_ + "vtu/");
        std::ofstream o(dir_ + "vtu/" + vtuName(f), std::ios::out | std::ios::binary);
        if (!o) chimp_host::die("Unable to open " + dir_ + "vtu/" + vtuName(f));
        const bool bin = FMT == VTK::BINARY;
        unsigned offset = 0;
        auto next = [&offset](std::size_t nbytes) { const unsigned at = offset; offset += unsigned(nbytes) + sizeof(unsigned); return at; };
        o << "<?xml version=\"1.0\"?>" << std::endl;
        o << "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">" << std::endl;
        o << "  <UnstructuredGrid>" << std::endl;
        o << "    <Piece NumberOfPoints=\"" << mesh_.nPoints << "\" NumberOfCells=\"" << mesh_.nCells << "\">" << std::endl;
        o << "      <Points>" << std::endl;
        o << arrayTag("points", "Float32", 3, next(mesh_.points.size() * sizeof(float)));
        if (!bin) for (float v : mesh_.points) asciiValue(o, v);
        o << "</DataArray>" << std::endl;
        o << "      </Points>" << std::endl;
        o << "      <Cells>" << std::endl;
        const std::vector<int> *cellArrays[3] = {&mesh_.connectivity, &mesh_.offsets, &mesh_.types};
        const char *cellNames[3] = {"connectivity", "offsets", "types"};
        for (int k = 0; k < 3; ++k) {
            o << arrayTag(cellNames[k], "Int32", 1, next(cellArrays[k]->size() * sizeof(int)));
            if (!bin) for (int v : *cellArrays[k]) asciiValue(o, v);
            o << "</DataArray>" << std::endl;
        }
        o << "      </Cells>" << std::endl;
        std::string scalars, vectors;
        for (const Array &a : f.arrays) {
            std::string &list = a.nComp > 1 ? vectors : scalars;
            list += (list.empty() ? "" : ", ") + a.name;
        }
        o << "      <CellData Scalars=\"" << scalars << "\" Vectors=\"" << vectors << "\">" << std::endl;
        for (const Array &a : f.arrays) {
            o << arrayTag(a.name, chimp_host::VtkTypeName<T>::name(), a.nComp, next(a.index.size() * sizeof(T)));
            if (!bin) for (int i : a.index) { if (i < 0) o << "0 "; else asciiValue(o, a.at(i)); }
            o << "</DataArray>" << std::endl;
        }
        o << "      </CellData>" << std::endl;
        o << "    </Piece>" << std::endl;
        o << "  </UnstructuredGrid>" << std::endl;
        if (bin) {
            o << "  <AppendedData encoding=\"raw\">" << std::endl;
            o << "_";
            raw(o, mesh_.points);
            for (int k = 0; k < 3; ++k) raw(o, *cellArrays[k]);
            std::vector<T> vals;
            for (const Array &a : f.arrays) {
                vals.resize(a.index.size());
                for (std::size_t k = 0; k < vals.size(); ++k) vals[k] = a.index[k] < 0 ? T(0) : a.at(a.index[k]);
                raw(o, vals);
            }
            o << "  </AppendedData>" << std::endl;
        }
        o << "</VTKFile>" << std::endl;
    }

    // header + one 100-byte record per rank + footer; every rank writes its own part at a computed position
    void writePvtu(const File &f, double time) const
    {
        std::ostringstream name;
        name << std::setfill('0') << f.name << "_" << std::setw(7) << f.nWrite << ".pvtu";
        const std::time_t now = std::time(nullptr);
        const std::tm tm = *std::localtime(&now);
        std::ostringstream h;
        h << "<?xml version=\"1.0\"?>" << std::endl;
        h << "<!-- Created " << std::setfill('0') << std::setw(2) << tm.tm_mday << "." << std::setw(2) << tm.tm_mon + 1 << "." << std::setw(2)
          << tm.tm_year + 1900 << " " << std::setw(2) << tm.tm_hour << ":" << std::setw(2) << tm.tm_min << ":" << std::setw(2) << tm.tm_sec
          << " -->" << std::endl;
        h << std::setfill(' ') << "<!-- time = " << time << " s -->" << std::endl;
        h << "<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">" << std::endl;
        h << "  <PUnstructuredGrid GhostLevel=\"0\">" << std::endl;
        h << "    <PPoints>" << std::endl;
        h << "      <PDataArray type=\"Float32\" NumberOfComponents=\"3\" />" << std::endl;
        h << "    </PPoints>" << std::endl;
        h << "    <PCellData>" << std::endl;
        h << "      <PDataArray type=\"Int32\" Name=\"connectivity\" />" << std::endl;
        h << "      <PDataArray type=\"Int32\" Name=\"offsets\" />" << std::endl;
        h << "      <PDataArray type=\"Int32\" Name=\"types\" />" << std::endl;
        h << "    </PCellData>" << std::endl;
        h << "    <PCellData>" << std::endl;
        for (const Array &a : f.arrays)
            h << "      <PDataArray Name=\"" << a.name << "\" type=\"" << chimp_host::VtkTypeName<T>::name() << "\" NumberOfComponents=\"" << a.nComp
              << "\" />" << std::endl;
        h << "    </PCellData>" << std::endl;
        const std::string header = h.str();
        constexpr int recordBytes = 100;
        char record[recordBytes + 1];
        std::snprintf(record, sizeof(record), "%-99s\n", ("    <Piece Source=\"vtu/" + vtuName(f) + "\" />").c_str());
        makeDir(dir_);
        const std::string path = dir_ + name.str();
        const int fd = ::open(path.c_str(), O_WRONLY | O_CREAT, 0644);
        if (fd < 0) chimp_host::die("Unable to open " + path);
        bool ok = true;
        if (rank_ == 0) ok = ok && ::pwrite(fd, header.data(), header.size(), 0) == ssize_t(header.size());
        ok = ok && ::pwrite(fd, record, recordBytes, off_t(header.size() + std::size_t(rank_) * recordBytes)) == recordBytes;
        if (rank_ == maxRank_) {
            const std::string footer = "  </PUnstructuredGrid>\n</VTKFile>\n";
            const off_t end = off_t(header.size() + std::size_t(maxRank_ + 1) * recordBytes);
            ok = ok && ::pwrite(fd, footer.data(), footer.size(), end) == ssize_t(footer.size());
            ok = ok && ::ftruncate(fd, end + off_t(footer.size())) == 0;
        }
        ::close(fd);
        if (!ok) chimp_host::die("write to " + path + " failed");
    }

    chimp_host::CellMesh mesh_;
    std::string dir_;
    int rank_, maxRank_;
    const std::vector<int> *nodes_ = nullptr;
    std::vector<File> files_;
};

#endif
