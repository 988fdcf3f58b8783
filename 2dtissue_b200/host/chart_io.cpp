#include "chart_io.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace t2dhost {

namespace {

bool ends_with(const std::string& s, const std::string& suf)
{
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

// OFF reader: "OFF" / "V F E" / V vertex lines / F lines "3 a b c".  Coordinates go through float, exactly like
// pmp::read_off with pmp::Scalar = float (MeshCartographyLib/pmp-library/src/pmp/types.h:16-20), then widen.
void read_off(const std::string& path, std::vector<double>& xyz, std::vector<int32_t>& faces, int& V, int& F)
{
    std::ifstream in(path);
    if (!in) throw std::runtime_error("Failed to open " + path + " (run the reference's chart setup first: it writes <stem>_uv.off and <stem>_open.off)");
    std::string tok;
    in >> tok;
    if (tok != "OFF") throw std::runtime_error(path + " is not an OFF file");
    long e = 0;
    in >> V >> F >> e;
    xyz.resize(3 * (size_t)V);
    for (size_t i = 0; i < 3 * (size_t)V; ++i) {
        std::string t;
        in >> t;
        xyz[i] = (double)std::strtof(t.c_str(), nullptr);
    }
    faces.resize(3 * (size_t)F);
    for (int f = 0; f < F; ++f) {
        int k = 0;
        in >> k;
        if (k != 3) throw std::runtime_error(path + ": only triangle meshes are supported");
        for (int j = 0; j < 3; ++j) in >> faces[3 * (size_t)f + j];
    }
    if (!in) throw std::runtime_error(path + ": truncated OFF file");
}

Chart load_t2dchart(const std::string& path)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("Failed to open " + path);
    char magic[8];
    uint32_t hdr[4];
    char pad[8];
    in.read(magic, 8);
    in.read(reinterpret_cast<char*>(hdr), sizeof(hdr));
    in.read(pad, 8);
    if (!in || std::memcmp(magic, "T2DCHART", 8) != 0 || hdr[0] != 1) throw std::runtime_error(path + " is not a T2DCHART v1 file");
    Chart c;
    c.V = (int)hdr[1];
    c.F = (int)hdr[2];
    const size_t P = hdr[3];
    c.uv.resize(2 * (size_t)c.V);
    c.x3d.resize(3 * (size_t)c.V);
    c.faces.resize(3 * (size_t)c.F);
    c.polygon.resize(2 * P);
    in.read(reinterpret_cast<char*>(c.uv.data()), sizeof(double) * c.uv.size());
    in.read(reinterpret_cast<char*>(c.x3d.data()), sizeof(double) * c.x3d.size());
    in.read(reinterpret_cast<char*>(c.faces.data()), sizeof(int32_t) * c.faces.size());
    in.read(reinterpret_cast<char*>(c.polygon.data()), sizeof(double) * c.polygon.size());
    if (!in) throw std::runtime_error(path + ": truncated chart file");
    return c;
}

}  // namespace

Chart load_chart(const std::string& mesh_path)
{
    if (ends_with(mesh_path, ".t2dchart")) return load_t2dchart(mesh_path);
    if (!ends_with(mesh_path, ".off")) throw std::runtime_error("mesh path must end in .off or .t2dchart: " + mesh_path);
    const std::string stem = mesh_path.substr(0, mesh_path.size() - 4);
    Chart c;
    std::vector<double> uv3, x3;
    std::vector<int32_t> f_uv, f_open;
    int V1, F1, V2, F2;
    read_off(stem + "_uv.off", uv3, f_uv, V1, F1);       // loadMeshFaces(mesh_UV_path, face_UV), 2DTissue.cpp:107
    read_off(stem + "_open.off", x3, f_open, V2, F2);    // vertice_3D of the cut-open mesh
    if (V1 != V2 || F1 != F2) throw std::runtime_error("UV and open meshes do not match: " + stem);
    c.V = V1;
    c.F = F1;
    c.uv.resize(2 * (size_t)V1);
    for (int v = 0; v < V1; ++v) {
        c.uv[2 * (size_t)v] = uv3[3 * (size_t)v];
        c.uv[2 * (size_t)v + 1] = uv3[3 * (size_t)v + 1];
    }
    c.x3d = x3;
    c.faces = f_uv;
    return c;
}

void save_t2dchart(const std::string& path, const Chart& c)
{
    std::ofstream out(path, std::ios::binary);
    if (!out) throw std::runtime_error("cannot write " + path);
    const uint32_t hdr[4] = {1u, (uint32_t)c.V, (uint32_t)c.F, (uint32_t)(c.polygon.size() / 2)};
    const char pad[8] = {0};
    out.write("T2DCHART", 8);
    out.write(reinterpret_cast<const char*>(hdr), sizeof(hdr));
    out.write(pad, 8);
    out.write(reinterpret_cast<const char*>(c.uv.data()), sizeof(double) * c.uv.size());
    out.write(reinterpret_cast<const char*>(c.x3d.data()), sizeof(double) * c.x3d.size());
    out.write(reinterpret_cast<const char*>(c.faces.data()), sizeof(int32_t) * c.faces.size());
    out.write(reinterpret_cast<const char*>(c.polygon.data()), sizeof(double) * c.polygon.size());
}

}  // namespace t2dhost
