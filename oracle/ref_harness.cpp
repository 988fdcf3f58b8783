/* oracle/ref_harness.cpp — TEST INFRASTRUCTURE, not product code.
 *
 * A thin C-ABI harness around the UNMODIFIED reference sources (compiled where they lie under
 * /root/reference by oracle/Makefile, output only into oracle/_ref/).  It drives the reference's own
 * classes for the hot path so that
 *   - tests/golden/ fixtures can be generated from the real reference (tools/make_golden.py),
 *   - the C restatement oracle/t2d_oracle.c can be pinned against it,
 *   - bench.py --impl reference can time the reference's own CPU step on the GPU box's host cores.
 * Nothing under 2dtissue_b200/ may link or load this.
 *
 * Reference entry points driven (file:line under /root/reference):
 *   SurfaceParametrization::create_uv_surface      MeshCartographyLib/src/SurfaceParametrization/SurfaceParametrization.cpp:81
 *   CachedGeodesicDistanceHelper::get_mesh_distance_matrix  MeshCartographyLib/src/GeodesicDistance/CachedGeodesicDistanceHelper.cpp:29
 *   Locomotion::simulate_flight                    src/simulation/Locomotion.cpp:57
 *   ForceHelper::calculate_forces_between_particles  src/simulation/Locomotion/ForceHelper.cpp:34
 *   OrientationHelper::calculate_average_n_within_distance  src/simulation/Locomotion/OrientationHelper.cpp:29
 *   LinearAlgebra::angles_to_unit_vectors          src/simulation/LinearAlgebra.cpp:25
 *   EuclideanTiling::diagonal_seam_edges_square_border  src/simulation/Locomotion/EuclideanTiling.cpp:31
 *   CellHelper::get_r3d                            src/simulation/CellHelper.cpp:71
 *   Validation::find_inside_uv_vertices_id / checkForInvalidValues  src/simulation/Validation.cpp:24,48
 *
 * Non-arithmetic shims: Tessellation::{left,right,up,down}_border are public but never
 * filled by the reference (TessellationHelper.h:17-25) — the harness fills them with the four sides of
 * the unit square as 2-point segments, in the order the reference iterates them (EuclideanTiling.cpp:79).
 *
 * Determinism shim: CellHelper.cpp:114 converts a 2-coefficient row block into an Eigen::Vector3d, which
 * (asserts compiled out in Release) reads r_UV.data()[i + 2N] — N doubles past the end of the heap block
 * (SURVEY.md §0).  The library is linked with -Wl,--wrap=malloc,--wrap=realloc and the wrappers below pad
 * every allocation by half its size and zero the pad, so that over-read always sees 0.0 (z := 0, the value
 * the restatement assumes) instead of heap garbage.  No reference source line or arithmetic changes.
 *
 * Host array layout == Eigen column-major: uv = N x's then N y's; r3d = N x, N y, N z.
 */
#include <memory>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include <filesystem>
#include <stdexcept>

#include "SurfaceParametrization/SurfaceParametrization.h"
#include "SurfaceParametrization/TessellationHelper.h"
#include "GeodesicDistance/CachedGeodesicDistanceHelper.h"
#include "Locomotion.h"
#include "Locomotion/ForceHelper.h"
#include "Locomotion/OrientationHelper.h"
#include "Locomotion/EuclideanTiling.h"
#include "CellHelper.h"
#include "LinearAlgebra.h"
#include "Validation.h"

extern "C" {
void* __real_malloc(size_t);
void* __real_realloc(void*, size_t);
static const size_t PAD_LIMIT = (size_t)64 << 20; /* only particle-sized blocks are padded (not the 178 MB table) */
void* __wrap_malloc(size_t n)
{
    if (n == 0 || n > PAD_LIMIT)
        return __real_malloc(n);
    const size_t pad = n / 2 + 64;
    char* p = (char*)__real_malloc(n + pad);
    if (p)
        std::memset(p + n, 0, pad);
    return p;
}
void* __wrap_realloc(void* q, size_t n)
{
    if (n == 0 || n > PAD_LIMIT)
        return __real_realloc(q, n);
    const size_t pad = n / 2 + 64;
    char* p = (char*)__real_realloc(q, n + pad);
    if (p)
        std::memset(p + n, 0, pad);
    return p;
}
}

namespace {

struct RefWorld {
    SurfaceParametrization sp;
    Tessellation tess{sp};
    Eigen::MatrixXd vertice_UV, vertice_3D, distance_matrix;
    Eigen::MatrixXi face_UV, face_3D;
    std::string mesh_path, mesh_uv_path;
    bool have_chart = false, have_table = false;
    std::string err;

    void fill_borders()
    {
        tess.left_border = {Point_2_eigen(0, 0), Point_2_eigen(0, 1)};
        tess.right_border = {Point_2_eigen(1, 0), Point_2_eigen(1, 1)};
        tess.up_border = {Point_2_eigen(0, 1), Point_2_eigen(1, 1)};
        tess.down_border = {Point_2_eigen(0, 0), Point_2_eigen(1, 0)};
    }
};

RefWorld* W = nullptr;
RefWorld& world()
{
    if (!W)
        W = new RefWorld();
    return *W;
}

using MatN2 = Eigen::Matrix<double, Eigen::Dynamic, 2>;

} // namespace

extern "C" {

const char* t2dref_last_error() { return world().err.c_str(); }

/* Run the reference's chart setup on a mesh that lives in <MeshCartographyLib_SOURCE_DIR>/meshes. */
int t2dref_chart_from_mesh(const char* mesh_path)
{
    RefWorld& w = world();
    try
    {
        w.mesh_path = mesh_path;
        std::tie(std::ignore, w.vertice_UV, w.vertice_3D, w.mesh_uv_path) = w.sp.create_uv_surface(w.mesh_path, 0);
        loadMeshFaces(w.mesh_uv_path, w.face_UV);
        loadMeshFaces(w.mesh_path, w.face_3D);
        w.fill_borders();
        w.have_chart = true;
    }
    catch (const std::exception& e)
    {
        w.err = e.what();
        return -1;
    }
    return 0;
}

/* Build (or load from the reference's own CSV cache) the vertex-distance table exactly as 2DTissue.cpp:102-105. */
int t2dref_table_build()
{
    RefWorld& w = world();
    if (!w.have_chart)
        return -2;
    try
    {
        fs::path path(w.mesh_path);
        fs::path mesh_open = path.parent_path() / (path.stem().string() + "_open.off");
        CachedGeodesicDistanceHelper helper_3D = CachedGeodesicDistanceHelper(mesh_open);
        GeodesicDistanceHelperInterface& geodesic_distance_helper_3D = helper_3D;
        w.distance_matrix = geodesic_distance_helper_3D.get_mesh_distance_matrix();
        w.have_table = true;
    }
    catch (const std::exception& e)
    {
        w.err = e.what();
        return -1;
    }
    return 0;
}

int t2dref_chart_sizes(int* V, int* F, int* P)
{
    RefWorld& w = world();
    if (!w.have_chart)
        return -2;
    *V = (int)w.vertice_UV.rows();
    *F = (int)w.face_UV.rows();
    *P = (int)w.sp.polygon.size();
    return 0;
}

/* Row-major exports: uv[V][2], x3d[V][3], faces[F][3], polygon[P][2]. */
int t2dref_chart_export(double* uv, double* x3d, int* faces, double* polygon)
{
    RefWorld& w = world();
    if (!w.have_chart)
        return -2;
    const int V = (int)w.vertice_UV.rows(), F = (int)w.face_UV.rows(), P = (int)w.sp.polygon.size();
    for (int i = 0; i < V; ++i)
    {
        uv[2 * i] = w.vertice_UV(i, 0);
        uv[2 * i + 1] = w.vertice_UV(i, 1);
        for (int c = 0; c < 3; ++c)
            x3d[3 * i + c] = w.vertice_3D(i, c);
    }
    for (int f = 0; f < F; ++f)
        for (int c = 0; c < 3; ++c)
            faces[3 * f + c] = w.face_UV(f, c);
    for (int p = 0; p < P; ++p)
    {
        polygon[2 * p] = w.sp.polygon[p][0];
        polygon[2 * p + 1] = w.sp.polygon[p][1];
    }
    return 0;
}

/* Feed a chart from a fixture instead of running MeshCartographyLib (used on the GPU box, where
 * /root/reference does not exist).  Only public members of the reference classes are written. */
int t2dref_chart_import(int V, int F, int P, const double* uv, const double* x3d, const int* faces, const double* polygon)
{
    RefWorld& w = world();
    w.vertice_UV.resize(V, 3);
    w.vertice_3D.resize(V, 3);
    w.face_UV.resize(F, 3);
    w.face_3D.resize(F, 3);
    for (int i = 0; i < V; ++i)
    {
        w.vertice_UV(i, 0) = uv[2 * i];
        w.vertice_UV(i, 1) = uv[2 * i + 1];
        w.vertice_UV(i, 2) = 0;
        for (int c = 0; c < 3; ++c)
            w.vertice_3D(i, c) = x3d[3 * i + c];
    }
    for (int f = 0; f < F; ++f)
        for (int c = 0; c < 3; ++c)
            w.face_UV(f, c) = w.face_3D(f, c) = faces[3 * f + c];
    w.sp.polygon.clear();
    for (int p = 0; p < P; ++p)
        w.sp.polygon.push_back(Point_2_eigen(polygon[2 * p], polygon[2 * p + 1]));
    w.fill_borders();
    w.have_chart = true;
    return 0;
}

/* Row-major V*V export/import of the table (symmetric, so the order is immaterial). */
int t2dref_table_export(double* D)
{
    RefWorld& w = world();
    if (!w.have_table)
        return -2;
    const int V = (int)w.distance_matrix.rows();
    for (int i = 0; i < V; ++i)
        for (int j = 0; j < V; ++j)
            D[(size_t)i * V + j] = w.distance_matrix(i, j);
    return 0;
}

int t2dref_table_import_u8(int V, const unsigned char* D)
{
    RefWorld& w = world();
    w.distance_matrix.resize(V, V);
    for (int i = 0; i < V; ++i)
        for (int j = 0; j < V; ++j)
            w.distance_matrix(i, j) = (double)D[(size_t)i * V + j];
    w.have_table = true;
    return 0;
}

int t2dref_table_import_f64(int V, const double* D)
{
    RefWorld& w = world();
    w.distance_matrix.resize(V, V);
    for (int i = 0; i < V; ++i)
        for (int j = 0; j < V; ++j)
            w.distance_matrix(i, j) = D[(size_t)i * V + j];
    w.have_table = true;
    return 0;
}

/* SurfaceParametrization::check_point_in_polygon on a batch. */
int t2dref_inside(int N, const double* uv, int* inside)
{
    RefWorld& w = world();
    for (int i = 0; i < N; ++i)
        inside[i] = w.sp.check_point_in_polygon(Point_2_eigen(uv[i], uv[N + i])) ? 1 : 0;
    return 0;
}

/* LinearAlgebra::angles_to_unit_vectors; out = N cos then N sin (column-major N x 2). */
int t2dref_angles_to_unit_vectors(int N, const int* n, double* out)
{
    LinearAlgebra la;
    Eigen::VectorXi nn = Eigen::Map<const Eigen::VectorXi>(n, N);
    MatN2 v = la.angles_to_unit_vectors(nn);
    std::memcpy(out, v.data(), sizeof(double) * 2 * N);
    return 0;
}

/* CellHelper::get_r3d */
int t2dref_get_r3d(int N, const double* uv, double* r3d, int* vid)
{
    RefWorld& w = world();
    if (!w.have_chart)
        return -2;
    MatN2 r_UV = Eigen::Map<const MatN2>(uv, N, 2);
    Eigen::MatrixXd r_3D(N, 3);
    Eigen::VectorXi n = Eigen::VectorXi::Zero(N);
    CellHelper ch(N, w.face_UV, w.face_3D, w.vertice_UV, w.vertice_3D, r_UV, r_3D, n);
    auto [pts, ids] = ch.get_r3d();
    std::memcpy(r3d, pts.data(), sizeof(double) * 3 * N);
    std::memcpy(vid, ids.data(), sizeof(int) * N);
    return 0;
}

/* EuclideanTiling::diagonal_seam_edges_square_border on (uv_old, uv, n), all in/out. */
int t2dref_tiling(int N, double* uv_old, double* uv, int* n)
{
    RefWorld& w = world();
    if (!w.have_chart)
        return -2;
    MatN2 r_UV = Eigen::Map<const MatN2>(uv, N, 2);
    MatN2 r_UV_old = Eigen::Map<const MatN2>(uv_old, N, 2);
    Eigen::VectorXi nn = Eigen::Map<const Eigen::VectorXi>(n, N);
    EuclideanTiling tiling(w.sp, w.tess, r_UV, r_UV_old, nn);
    tiling.diagonal_seam_edges_square_border();
    std::memcpy(uv, r_UV.data(), sizeof(double) * 2 * N);
    std::memcpy(uv_old, r_UV_old.data(), sizeof(double) * 2 * N);
    std::memcpy(n, nn.data(), sizeof(int) * N);
    return 0;
}

/* ForceHelper + OrientationHelper on a caller-built dist_length (row-major N*N; symmetric in all uses). */
int t2dref_force_orientation(int N, const double* uv, int* n, const double* dist_length_rm, double k, double sigma,
                             double* F_out)
{
    MatN2 r_UV = Eigen::Map<const MatN2>(uv, N, 2);
    Eigen::MatrixXd dist_length(N, N);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j)
            dist_length(i, j) = dist_length_rm[(size_t)i * N + j];
    std::vector<Eigen::MatrixXd> dist_vect(2);
    dist_vect[0].resize(N, N);
    dist_vect[1].resize(N, N);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j)
        {
            dist_vect[0](i, j) = r_UV(i, 0) - r_UV(j, 0);
            dist_vect[1](i, j) = r_UV(i, 1) - r_UV(j, 1);
        }
    MatN2 F(N, 2);
    ForceHelper fh(F, k, sigma, 1.0, 0.75, dist_length, dist_vect);
    fh.calculate_forces_between_particles();
    Eigen::VectorXi nn = Eigen::Map<const Eigen::VectorXi>(n, N);
    OrientationHelper oh(dist_vect, dist_length, nn, sigma);
    oh.calculate_average_n_within_distance();
    std::memcpy(F_out, F.data(), sizeof(double) * 2 * N);
    std::memcpy(n, nn.data(), sizeof(int) * N);
    return 0;
}

/* One full timestep == the body of _2DTissue::perform_particle_simulation (2DTissue.cpp:222-252).
 *
 * mode 0 (table): Locomotion::simulate_flight() unmodified, i.e. dist_length(i,j)=D(vid_i,vid_j).
 * mode 1 (euclid, extension defined in SURVEY.md App. A): dist_length(i,j)=||X_i-X_j|| on the 3-D
 *   positions of the PREVIOUS projection (r3d in), then the reference's ForceHelper / OrientationHelper
 *   unchanged and simulate_flight lines 71-84 restated.
 * eta: optional per-particle noise in degrees, applied where OrientationHelper.cpp:70 applies it.
 * In/out: uv (2N), n (N), vid (N), r3d (3N).  Out: rdot (2N), color (N), F (2N, may be NULL).
 * Returns fault bitmask: 1 lost particle, 2 NaN/Inf (Validation.cpp:40-72), <0 harness error.
 */
int t2dref_step(int N, double* uv, int* n, int* vid, double* r3d, double* rdot, int* color, double* F_out, double v0,
                double k, double sigma, double step_size, const double* eta, int mode)
{
    RefWorld& w = world();
    if (!w.have_chart || (mode == 0 && !w.have_table))
        return -2;
    try
    {
        MatN2 r_UV = Eigen::Map<const MatN2>(uv, N, 2);
        MatN2 r_UV_old = r_UV;
        MatN2 r_dot(N, 2);
        Eigen::VectorXi nn = Eigen::Map<const Eigen::VectorXi>(n, N);
        std::vector<int> vertices_3D_active(vid, vid + N);
        Eigen::MatrixXd dist_length = Eigen::MatrixXd::Zero(N, N);
        Eigen::MatrixXd r_3D = Eigen::Map<const Eigen::MatrixXd>(r3d, N, 3);
        Eigen::MatrixXd dummy_table(1, 1);
        MatN2 F_keep = MatN2::Zero(N, 2);

        if (mode == 0)
        {
            Locomotion locomotion(r_UV, r_UV_old, r_dot, nn, vertices_3D_active, w.distance_matrix, dist_length, v0, k,
                                  sigma, 1.0, 1.0, 0.75, step_size, std::make_unique<LinearAlgebra>());
            locomotion.simulate_flight();
            if (F_out)
            {
                /* F_track is private; recompute it with the reference's own ForceHelper for export only. */
                std::vector<Eigen::MatrixXd> dv(2);
                dv[0].resize(N, N);
                dv[1].resize(N, N);
                for (int i = 0; i < N; ++i)
                    for (int j = 0; j < N; ++j)
                    {
                        dv[0](i, j) = r_UV_old(i, 0) - r_UV_old(j, 0);
                        dv[1](i, j) = r_UV_old(i, 1) - r_UV_old(j, 1);
                    }
                ForceHelper fh(F_keep, k, sigma, 1.0, 0.75, dist_length, dv);
                fh.calculate_forces_between_particles();
            }
        }
        else
        {
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j)
                {
                    double dx = r_3D(i, 0) - r_3D(j, 0), dy = r_3D(i, 1) - r_3D(j, 1), dz = r_3D(i, 2) - r_3D(j, 2);
                    dist_length(i, j) = (i == j) ? 0.0 : std::sqrt(dx * dx + dy * dy + dz * dz);
                }
            std::vector<Eigen::MatrixXd> dist_vect(2);
            dist_vect[0].resize(N, N);
            dist_vect[1].resize(N, N);
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j)
                {
                    dist_vect[0](i, j) = r_UV(i, 0) - r_UV(j, 0);
                    dist_vect[1](i, j) = r_UV(i, 1) - r_UV(j, 1);
                }
            ForceHelper fh(F_keep, k, sigma, 1.0, 0.75, dist_length, dist_vect);
            fh.calculate_forces_between_particles();
            /* Locomotion.cpp:71-84 restated */
            LinearAlgebra la;
            Eigen::VectorXd abs_F = F_keep.rowwise().norm();
            MatN2 n_vec = la.angles_to_unit_vectors(nn);
            abs_F = abs_F.array() + v0;
            r_dot = n_vec.array().colwise() * abs_F.array();
            r_UV += r_dot * step_size;
            OrientationHelper oh(dist_vect, dist_length, nn, sigma);
            oh.calculate_average_n_within_distance();
        }

        if (eta)
            for (int i = 0; i < N; ++i)
                nn(i) += eta[i]; /* int += double, as OrientationHelper.cpp:70 */

        EuclideanTiling tiling(w.sp, w.tess, r_UV, r_UV_old, nn);
        tiling.diagonal_seam_edges_square_border();

        CellHelper ch(N, w.face_UV, w.face_3D, w.vertice_UV, w.vertice_3D, r_UV, r_3D, nn);
        auto [new_r_3D, new_vertices_3D_active] = ch.get_r3d();

        Validation validation(w.sp);
        int fault = 0;
        if ((int)validation.find_inside_uv_vertices_id(r_UV).size() != N)
            fault |= 1;
        if (validation.checkForInvalidValues(r_UV))
            fault |= 2;

        /* _2DTissue::count_particle_neighbors (2DTissue.cpp:254-268) restated — it is a private member. */
        for (int i = 0; i < N; ++i)
        {
            int c = 0;
            for (int j = 0; j < N; ++j)
                if (dist_length(i, j) != 0 && dist_length(i, j) <= 2.4 * sigma)
                    c += 1;
            color[i] = c;
        }

        std::memcpy(uv, r_UV.data(), sizeof(double) * 2 * N);
        std::memcpy(n, nn.data(), sizeof(int) * N);
        std::memcpy(vid, new_vertices_3D_active.data(), sizeof(int) * N);
        std::memcpy(r3d, new_r_3D.data(), sizeof(double) * 3 * N);
        std::memcpy(rdot, r_dot.data(), sizeof(double) * 2 * N);
        if (F_out)
            std::memcpy(F_out, F_keep.data(), sizeof(double) * 2 * N);
        return fault;
    }
    catch (const std::exception& e)
    {
        w.err = e.what();
        return -1;
    }
}

} // extern "C"
