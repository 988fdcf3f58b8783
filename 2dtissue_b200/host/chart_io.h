// chart_io.h — the UV chart of a cut-open mesh as the reference's host-side setup leaves it for the stepping loop
// (/root/reference/src/simulation/2DTissue.cpp:85-107: vertice_UV, vertice_3D, face_UV).  The chart is PRODUCED by the
// reference's MeshCartographyLib (host-side surface, out of scope here); this file only reads what that setup writes:
//   <stem>_uv.off + <stem>_open.off   (SurfaceParametrization.cpp:118-128; text, "%.10f"; parsed as float like pmp::Scalar)
//   or a binary .t2dchart             (2dtissue_b200/chart.py, same layout)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace t2dhost {

struct Chart {
    int V = 0, F = 0;
    std::vector<double> uv;      // [V][2]
    std::vector<double> x3d;     // [V][3]
    std::vector<int32_t> faces;  // [F][3]
    std::vector<double> polygon; // [P][2] border polygon (unused by the step; kept for round trips)
};

// mesh_path: "<dir>/<stem>.off" (then <stem>_uv.off and <stem>_open.off next to it are read), or a .t2dchart file.
// Throws std::runtime_error with the reference's own wording when a file is missing.
Chart load_chart(const std::string& mesh_path);
void save_t2dchart(const std::string& path, const Chart& c);

}  // namespace t2dhost
