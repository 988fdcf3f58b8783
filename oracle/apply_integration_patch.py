#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Applies the patch of INTEGRATION.md §2 to a scratch copy of the reference's `2DTissue.{h,cpp}`
(written to oracle/_ref/patched/, a git-ignored build directory — no reference source enters the repository) so that the
reference binary itself runs its step on lib2dtissue_b200.so:

  * `2DTissue.h`: `#include "t2d.h"`, one new member `t2d_ctx* gpu`, and a declaration for the renamed stock body;
  * constructor: `t2d_create` from the chart and the distance matrix the reference has just built;
  * `start()`: `t2d_set_particles` (upload + initial get_r3d on the GPU);
  * `perform_particle_simulation()`: the body becomes ONE call, `t2d_step_host`; the stock body is kept under the name
    `perform_particle_simulation_stock()` and, when T2D_INTEGRATION_CHECK is set, is run from the same state right after
    the library so that every step of the run is compared (uv, heading, nearest vertex, 3-D position, velocity, colour).

The anchors are function names and member names of the reference; the inserted text is this repository's.
usage: apply_integration_patch.py REF_DIR OUT_DIR
"""
import os
import sys

ref, out = sys.argv[1], sys.argv[2]
os.makedirs(out, exist_ok=True)
h = open(os.path.join(ref, "src/simulation/2DTissue.h"), encoding="utf-8").read()
c = open(os.path.join(ref, "src/simulation/2DTissue.cpp"), encoding="utf-8").read()


def once(text, anchor, repl, what):
    if text.count(anchor) != 1:
        raise SystemExit("integration patch: anchor for %s found %d times" % (what, text.count(anchor)))
    return text.replace(anchor, repl)


# ---- header -------------------------------------------------------------------------------------------------------
h = once(h, "    Locomotion locomotion;", "    t2d_ctx* gpu = nullptr;   // INTEGRATION.md: owns the device copy of chart, table and particle state\n"
            "    void perform_particle_simulation_stock();\n    Locomotion locomotion;", "the new member")
h = '#include "t2d.h"\n' + h

# ---- constructor: after the last statement of the reference's constructor ------------------------------------------
CTOR = r'''
    // ---- INTEGRATION.md §2: hand the chart and the table to the library ----
    {
        Eigen::Matrix<double, Eigen::Dynamic, 2, Eigen::RowMajor> uv_rm = vertice_UV.leftCols<2>();
        Eigen::Matrix<double, Eigen::Dynamic, 3, Eigen::RowMajor> x3_rm = vertice_3D;
        Eigen::Matrix<int, Eigen::Dynamic, 3, Eigen::RowMajor> f_rm = face_UV;
        Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> D_rm = distance_matrix;
        t2d_mesh mesh{(int32_t)vertice_UV.rows(), (int32_t)face_UV.rows(), uv_rm.data(), x3_rm.data(), f_rm.data()};
        t2d_table table{(int32_t)distance_matrix.rows(), T2D_TABLE_DENSE_F64, D_rm.data()};
        t2d_params prm{v0, k, σ, step_size, /*eta*/ 0.0, /*color_factor*/ 2.4, /*seed*/ 0,
                       T2D_NEIGH_TABLE, T2D_PRECISION_FP64, particle_count, T2D_LIFT_REFERENCE};
        if (t2d_create(&mesh, &table, &prm, /*device*/ 0, &gpu) != 0) throw std::runtime_error(t2d_last_error(nullptr));
    }
'''
c = once(c, "    dist_length = Eigen::MatrixXd::Zero(particle_count, particle_count);\n",
         "    dist_length = Eigen::MatrixXd::Zero(particle_count, particle_count);\n" + CTOR, "the constructor")

# ---- start(): upload + initial projection on the GPU, checked against the stock get_r3d -----------------------------
START = r'''
    // ---- INTEGRATION.md §2 ----
    if (t2d_set_particles(gpu, particle_count, r_UV.data(), n.data(), nullptr) != 0) throw std::runtime_error(t2d_last_error(gpu));
    if (std::getenv("T2D_INTEGRATION_CHECK")) {
        std::vector<int> va(particle_count);
        Eigen::MatrixXd r3(particle_count, 3);
        t2d_download(gpu, nullptr, nullptr, va.data(), r3.data(), nullptr, nullptr, nullptr);
        int bad = 0;
        for (int i = 0; i < particle_count; ++i) bad += va[i] != vertices_3D_active[i];
        std::printf("T2D_CHECK start: vid mismatches %d, max |r3d diff| %.3g\n", bad, (r3 - r_3D).cwiseAbs().maxCoeff());
    }
'''
c = once(c, "    std::tie(r_3D, vertices_3D_active) = cell_helper.get_r3d();\n",
         "    std::tie(r_3D, vertices_3D_active) = cell_helper.get_r3d();\n" + START, "start()")

# ---- the step -------------------------------------------------------------------------------------------------------
c = once(c, "void _2DTissue::perform_particle_simulation()", "void _2DTissue::perform_particle_simulation_stock()", "the stock body")
STEP = r'''
// ---- INTEGRATION.md §2: the whole body of the step is one call into lib2dtissue_b200.so ----
void _2DTissue::perform_particle_simulation()
{
    const bool check = std::getenv("T2D_INTEGRATION_CHECK") != nullptr;
    const auto uv0 = r_UV;
    const auto uvold0 = r_UV_old;
    const auto n0 = n;
    const auto r3d0 = r_3D;
    const auto va0 = vertices_3D_active;
    int fault = t2d_step_host(gpu, particle_count, r_UV.data(), n.data(), vertices_3D_active.data(), r_3D.data(),
                              r_dot.data(), particles_color.data());
    if (fault < 0) throw std::runtime_error(t2d_last_error(gpu));
    if (fault & T2D_FAULT_LOST) throw std::runtime_error("We lost particles after getting the original UV mesh coord");
    if (fault & T2D_FAULT_NONFINITE) std::exit(1);
    double obs[T2D_OBS_LEN];
    t2d_observables(gpu, obs);
    v_order(current_step) = obs[T2D_OBS_PHI];
    if (!check) return;
    // the stock body from the same state, then compare and continue from the stock result
    const auto uv_g = r_UV;
    const auto n_g = n;
    const auto r3d_g = r_3D;
    const auto va_g = vertices_3D_active;
    const auto rdot_g = r_dot;
    const auto col_g = particles_color;
    r_UV = uv0;
    r_UV_old = uvold0;
    n = n0;
    r_3D = r3d0;
    vertices_3D_active = va0;
    std::fill(particles_color.begin(), particles_color.end(), 0);
    perform_particle_simulation_stock();
    int n_bad = 0, va_bad = 0, col_bad = 0;
    double duv = 0, dr3 = 0, drd = 0;
    for (int i = 0; i < particle_count; ++i) {
        col_bad += col_g[i] != particles_color[i];
        drd = std::max(drd, std::max(std::abs(rdot_g(i, 0) - r_dot(i, 0)), std::abs(rdot_g(i, 1) - r_dot(i, 1))));
        if (n_g(i) != n(i)) {   // integer heading off a truncation tie (|angle - rint(angle)| < 1e-9): the seam logic may differ too
            ++n_bad;
            continue;
        }
        va_bad += va_g[i] != vertices_3D_active[i];
        duv = std::max(duv, std::max(std::abs(uv_g(i, 0) - r_UV(i, 0)), std::abs(uv_g(i, 1) - r_UV(i, 1))));
        for (int k3 = 0; k3 < 3; ++k3) dr3 = std::max(dr3, std::abs(r3d_g(i, k3) - r_3D(i, k3)));
    }
    std::printf("T2D_CHECK step %d: heading mismatches %d, vid mismatches %d, colour mismatches %d, max |uv diff| %.3g, "
                "max |r3d diff| %.3g, max |rdot diff| %.3g\n", current_step, n_bad, va_bad, col_bad, duv, dr3, drd);
}
'''
c = c + STEP
c = "#include <cstdio>\n#include <cstdlib>\n" + c
open(os.path.join(out, "2DTissue.h"), "w", encoding="utf-8").write(h)
open(os.path.join(out, "2DTissue.cpp"), "w", encoding="utf-8").write(c)
print("integration patch applied ->", out)
