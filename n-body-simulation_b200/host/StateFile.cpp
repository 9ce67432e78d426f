#include "StateFile.hpp"

#include <cstdio>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <vector>

namespace StateFile {

namespace {

struct Header {
    char magic[8];
    std::uint32_t version;
    std::uint32_t flags;
    std::uint64_t bodies;
    double time;
    unsigned char reserved[32];
};
static_assert(sizeof(Header) == kHeaderBytes, "state file header layout");

void readArray(std::ifstream &in, std::vector<double> &v, std::size_t n, const std::string &path) {
    v.resize(n);
    in.read(reinterpret_cast<char *>(v.data()), (std::streamsize) (n * sizeof(double)));
    if ((std::size_t) in.gcount() != n * sizeof(double)) throw std::invalid_argument("state file is truncated: " + path);
}

}  // namespace

bool isStateFile(const std::string &path) {
    std::ifstream in(path, std::ios::binary);
    char magic[8] = {};
    in.read(magic, 8);
    return in.gcount() == 8 && std::memcmp(magic, kMagic, 8) == 0;
}

void read(const std::string &path, SimulationData &d, double *time) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::invalid_argument("cannot open input file " + path);
    Header h;
    in.read(reinterpret_cast<char *>(&h), sizeof h);
    if (in.gcount() != (std::streamsize) sizeof h || std::memcmp(h.magic, kMagic, 8) != 0)
        throw std::invalid_argument("not a body-state file: " + path);
    if (h.version != kVersion) throw std::invalid_argument("unsupported state file version in " + path);
    const std::size_t n = (std::size_t) h.bodies;
    readArray(in, d.mass, n, path);
    readArray(in, d.positions_x, n, path);
    readArray(in, d.positions_y, n, path);
    readArray(in, d.positions_z, n, path);
    readArray(in, d.velocities_x, n, path);
    readArray(in, d.velocities_y, n, path);
    readArray(in, d.velocities_z, n, path);
    d.names.clear();
    d.body_classes.clear();
    if (h.flags & 1u) {
        d.names.reserve(n);
        d.body_classes.reserve(n);
        std::string name, cls;
        for (std::size_t i = 0; i < n; ++i) {
            if (!std::getline(in, name, '\0') || !std::getline(in, cls, '\0'))
                throw std::invalid_argument("state file name table is truncated: " + path);
            d.names.push_back(name);
            d.body_classes.push_back(cls);
        }
    }
    if (time) *time = h.time;
}

void write(const std::string &path, const SimulationData &d, double time) {
    const std::size_t n = d.mass.size();
    if (d.positions_x.size() != n || d.positions_y.size() != n || d.positions_z.size() != n ||
        d.velocities_x.size() != n || d.velocities_y.size() != n || d.velocities_z.size() != n)
        throw std::invalid_argument("state arrays differ in length");
    const double *arrays[7] = {d.mass.data(), d.positions_x.data(), d.positions_y.data(), d.positions_z.data(),
                               d.velocities_x.data(), d.velocities_y.data(), d.velocities_z.data()};
    writeArrays(path, n, arrays, d.names, d.body_classes, time);
}

void writeArrays(const std::string &path, std::size_t n, const double *const arrays[7],
                 const std::vector<std::string> &names, const std::vector<std::string> &classes, double time) {
    const bool table = names.size() == n && classes.size() == n && n > 0;
    Header h;
    std::memset(&h, 0, sizeof h);
    std::memcpy(h.magic, kMagic, 8);
    h.version = kVersion;
    h.flags = table ? 1u : 0u;
    h.bodies = n;
    h.time = time;
    const std::string tmp = path + ".tmp";
    {
        std::ofstream out(tmp, std::ios::binary | std::ios::trunc);
        if (!out) throw std::invalid_argument("cannot write state file " + path);
        out.write(reinterpret_cast<const char *>(&h), sizeof h);
        for (int a = 0; a < 7; ++a)
            out.write(reinterpret_cast<const char *>(arrays[a]), (std::streamsize) (n * sizeof(double)));
        if (table)
            for (std::size_t i = 0; i < n; ++i) {
                out.write(names[i].c_str(), (std::streamsize) names[i].size() + 1);
                out.write(classes[i].c_str(), (std::streamsize) classes[i].size() + 1);
            }
        if (!out) throw std::invalid_argument("write failed for state file " + path);
    }
    // a checkpoint replaces the previous one atomically
    if (std::rename(tmp.c_str(), path.c_str()) != 0) throw std::invalid_argument("cannot move state file into place: " + path);
}

}  // namespace StateFile
