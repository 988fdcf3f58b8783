#include "tissue.h"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <random>
#include <stdexcept>

namespace t2dhost {

Tissue2D::Tissue2D(bool save_data_, bool particle_innenleben, bool free_boundary, std::string mesh_path, int particle_count_,
                   int step_count_, double v0_, bool use_kafka, double k_, double k_next, double v0_next, double sigma_,
                   double mu, double r_adh, double k_adh, double step_size_, int map_cache_count, const Extensions& ext)
    : save_data(save_data_), particle_count(particle_count_), step_count(step_count_), v0(v0_), k(k_), sigma(sigma_),
      step_size(step_size_), ext_(ext)
{
    // accepted and ignored exactly like the reference
    (void)particle_innenleben; (void)free_boundary; (void)k_next; (void)v0_next; (void)mu; (void)r_adh; (void)k_adh;
    (void)map_cache_count;
    if (use_kafka) throw std::runtime_error("Kafka streaming is not part of this build (output sink, out of scope)");

    chart_ = load_chart(mesh_path);   // the reference's chart setup (2DTissue.cpp:85-107) stays host-side and upstream

    t2d_mesh mesh{chart_.V, chart_.F, chart_.uv.data(), chart_.x3d.data(), chart_.faces.data()};
    // table criterion: the reference's hop-count table (CachedGeodesicDistanceHelper, 2DTissue.cpp:102-105) is built on
    // the GPU from the chart's edge graph instead of being parsed from a 44 MB CSV
    t2d_table table{chart_.V, ext_.neigh_mode == T2D_NEIGH_TABLE ? T2D_TABLE_HOPS_FROM_MESH : T2D_TABLE_NONE, nullptr};
    // --table FILE: thresholded rows from the binary cache (replaces the reference's CSV cache, CachedGeodesicDistanceHelper.h:22-72)
    std::vector<int32_t> c_start, c_col;
    std::vector<unsigned char> c_val;
    t2d_table_csr csr{};
    if (ext_.neigh_mode == T2D_NEIGH_TABLE && !ext_.table_cache.empty()) {
        std::ifstream in(ext_.table_cache, std::ios::binary);
        char magic[8];
        int64_t hdr[2], item = 0;
        double radius = 0;
        in.read(magic, 8);
        in.read(reinterpret_cast<char*>(hdr), sizeof(hdr));
        in.read(reinterpret_cast<char*>(&radius), sizeof(radius));
        in.read(reinterpret_cast<char*>(&item), sizeof(item));
        if (!in || std::string(magic, 8) != std::string("T2DCSR1\0", 8) || hdr[0] != chart_.V || (item != 8 && item != 4 && item != 1))
            throw std::runtime_error("cannot read the table cache " + ext_.table_cache + " (missing, not T2DCSR1, or for another mesh)");
        c_start.resize((size_t)hdr[0] + 1);
        c_col.resize((size_t)hdr[1]);
        c_val.resize((size_t)hdr[1] * (size_t)item);
        in.read(reinterpret_cast<char*>(c_start.data()), sizeof(int32_t) * c_start.size());
        in.read(reinterpret_cast<char*>(c_col.data()), sizeof(int32_t) * c_col.size());
        in.read(reinterpret_cast<char*>(c_val.data()), (std::streamsize)c_val.size());
        if (!in) throw std::runtime_error("truncated table cache " + ext_.table_cache);
        csr = t2d_table_csr{hdr[1], c_start.data(), c_col.data(), c_val.data(), radius};
        table.kind = item == 8 ? T2D_TABLE_CSR_F64 : (item == 4 ? T2D_TABLE_CSR_F32 : T2D_TABLE_CSR_U8);
        table.data = &csr;
    }
    t2d_params prm{};
    prm.v0 = v0;
    prm.k = k;
    prm.sigma = sigma;
    prm.step_size = step_size;
    prm.eta = ext_.eta;
    prm.color_factor = 2.4;   // 2DTissue.cpp:262
    prm.seed = ext_.seed;
    prm.neigh_mode = ext_.neigh_mode;
    prm.lift_mode = ext_.lift_mode;
    prm.precision = ext_.precision;
    prm.capacity = particle_count > 0 ? particle_count : 1;
    if (t2d_create(&mesh, &table, &prm, ext_.device, &gpu) != 0) throw std::runtime_error(t2d_last_error(nullptr));

    v_order.assign((size_t)(step_count > 0 ? step_count : 0), 0.0);   // v_order = VectorXd::Zero(step_count)
}

Tissue2D::~Tissue2D() { t2d_destroy(gpu); }

void Tissue2D::check(int rc, const char* what)
{
    if (rc < 0) throw std::runtime_error(std::string(what) + ": " + t2d_last_error(gpu));
}

void Tissue2D::init_particle_position()
{
    std::random_device rd;
    std::mt19937 gen(ext_.seed ? (uint32_t)ext_.seed : rd());
    std::uniform_int_distribution<> dis_face(0, chart_.F - 1);
    std::uniform_int_distribution<> dis_angle(0, 359);
    const size_t N = (size_t)particle_count;
    for (size_t i = 0; i < N; ++i) {
        const int f = dis_face(gen);   // faces may repeat, as in the reference (App. C #12)
        const int32_t* fv = &chart_.faces[3 * (size_t)f];
        double cx = 0, cy = 0;         // get_face_gravity_center_coord: centroid of the UV face
        for (int q = 0; q < 3; ++q) {
            cx += chart_.uv[2 * (size_t)fv[q]];
            cy += chart_.uv[2 * (size_t)fv[q] + 1];
        }
        r_UV[i] = cx / 3.0;
        r_UV[N + i] = cy / 3.0;
        n[i] = dis_angle(gen);
    }
}

void Tissue2D::start()
{
    const size_t N = (size_t)particle_count;
    r_UV.assign(2 * N, 0.0);
    r_dot.assign(2 * N, 0.0);
    r_3D.assign(3 * N, 0.0);
    n.assign(N, 0);
    particles_color.assign(N, 0);
    vertices_3D_active.assign(N, 0);
    if (!ext_.load_state.empty()) {
        std::ifstream in(ext_.load_state, std::ios::binary);
        char magic[8];
        int64_t hdr[2];
        in.read(magic, 8);
        in.read(reinterpret_cast<char*>(hdr), sizeof(hdr));
        if (!in || std::string(magic, 8) != "T2DSTATE" || hdr[0] != (int64_t)N)
            throw std::runtime_error("cannot resume from " + ext_.load_state + " (missing, or particle count differs)");
        in.read(reinterpret_cast<char*>(r_UV.data()), sizeof(double) * 2 * N);
        in.read(reinterpret_cast<char*>(n.data()), sizeof(int32_t) * N);
        if (!in) throw std::runtime_error("truncated state file " + ext_.load_state);
        current_step = (int)hdr[1];
        t2d_set_step(gpu, hdr[1]);
    } else if (ext_.device_seed) {   // seeding + initial projection on the GPU: face centres like CellHelper::init_particle_position
        check(t2d_seed_particles(gpu, particle_count, ext_.seed ? ext_.seed : 1234, /*mode*/ 1, 0), "t2d_seed_particles");
        check(t2d_download(gpu, r_UV.data(), n.data(), vertices_3D_active.data(), r_3D.data(), nullptr, nullptr, nullptr), "t2d_download");
        return;
    } else {
        init_particle_position();
    }
    // upload + the initial CellHelper::get_r3d on the GPU (2DTissue.cpp:133)
    check(t2d_set_particles(gpu, particle_count, r_UV.data(), n.data(), nullptr), "t2d_set_particles");
    check(t2d_download(gpu, nullptr, nullptr, vertices_3D_active.data(), r_3D.data(), nullptr, nullptr, nullptr), "t2d_download");
}

System Tissue2D::update()
{
    if (!ext_.quiet) std::cout << "Step: " << current_step << "\n";

    // perform_particle_simulation() — the whole body runs on the GPU, state stays resident
    const int fault = t2d_step(gpu, 1);
    check(fault, "t2d_step");
    if (fault & T2D_FAULT_LOST)        // Validation::error_lost_particles, Validation.cpp:66-72
        throw std::runtime_error("We lost particles after getting the original UV mesh coord");
    if (fault & T2D_FAULT_NONFINITE) { // Validation::error_invalid_values, Validation.cpp:40-46
        std::cerr << "Invalid values (NaN or Inf) in the particle positions\n";
        std::exit(1);
    }
    if (fault & T2D_FAULT_WRAP_CAP) throw std::runtime_error("seam re-entry did not terminate (EuclideanTiling)");
    double obs[T2D_OBS_LEN];
    check(t2d_observables(gpu, obs), "t2d_observables");
    if (current_step < (int)v_order.size()) v_order[(size_t)current_step] = obs[T2D_OBS_PHI];
    if (!ext_.quiet) std::cout << "\n";

    System system;
    system.order_parameter = obs[T2D_OBS_PHI];
    if (ext_.export_particles || save_data) {
        check(t2d_download(gpu, r_UV.data(), n.data(), vertices_3D_active.data(), r_3D.data(), r_dot.data(),
                           particles_color.data(), nullptr), "t2d_download");
        const size_t N = (size_t)particle_count;
        if (ext_.export_particles) {
            system.particles.resize(N);
            for (size_t i = 0; i < N; ++i) {
                Particle p{};
                p.x_UV = r_UV[i];
                p.y_UV = r_UV[N + i];
                p.x_velocity_UV = r_dot[i];
                p.y_velocity_UV = r_dot[N + i];
                p.alignment_UV = n[i];
                p.x_3D = r_3D[i];
                p.y_3D = r_3D[N + i];
                p.z_3D = r_3D[2 * N + i];
                p.neighbor_count = particles_color[i];
                system.particles[i] = p;
            }
        }
    }
    current_step++;
    if (current_step >= step_count) finished = true;
    if (save_data) save_our_data();
    return system;
}

// ---- every-k cadence with the asynchronous export (include/t2d.h t2d_export_begin / t2d_export_wait) ----
System Tissue2D::collect_export(int slot)
{
    System system{};
    int32_t N = 0;
    int64_t step = 0;
    const double *uv = nullptr, *r3d = nullptr, *rdot = nullptr;
    const int32_t *h = nullptr, *vid = nullptr, *col = nullptr;
    check(t2d_export_wait(gpu, slot, &N, &step, &uv, &h, &vid, &r3d, &rdot, &col), "t2d_export_wait");
    const size_t n = (size_t)N;
    std::copy(uv, uv + 2 * n, r_UV.begin());
    std::copy(r3d, r3d + 3 * n, r_3D.begin());
    std::copy(rdot, rdot + 2 * n, r_dot.begin());
    std::copy(h, h + n, this->n.begin());
    std::copy(vid, vid + n, vertices_3D_active.begin());
    std::copy(col, col + n, particles_color.begin());
    const int idx = (int)step - 1;
    system.order_parameter = (idx >= 0 && idx < (int)v_order.size()) ? v_order[(size_t)idx] : 0.0;
    if (ext_.export_particles) {
        system.particles.resize(n);
        for (size_t i = 0; i < n; ++i) {
            Particle p{};
            p.x_UV = r_UV[i];
            p.y_UV = r_UV[n + i];
            p.x_velocity_UV = r_dot[i];
            p.y_velocity_UV = r_dot[n + i];
            p.alignment_UV = this->n[i];
            p.x_3D = r_3D[i];
            p.y_3D = r_3D[n + i];
            p.z_3D = r_3D[2 * n + i];
            p.neighbor_count = particles_color[i];
            system.particles[i] = p;
        }
    }
    if (save_data) {
        const int keep = current_step;
        current_step = (int)step;     // file names carry the step the snapshot belongs to
        save_our_data();
        current_step = keep;
    }
    return system;
}

System Tissue2D::update_block(int k)
{
    k = std::max(1, std::min(k, step_count - current_step));
    if (!ext_.quiet) std::cout << "Steps: " << current_step << " .. " << current_step + k - 1 << "\n";
    // the order parameter is a by-product of every step (cheap: one reduction); the state export is not
    for (int s = 0; s < k; ++s) {
        const int fault = t2d_step(gpu, 1);
        check(fault, "t2d_step");
        if (fault & T2D_FAULT_LOST) throw std::runtime_error("We lost particles after getting the original UV mesh coord");
        if (fault & T2D_FAULT_NONFINITE) {
            std::cerr << "Invalid values (NaN or Inf) in the particle positions\n";
            std::exit(1);
        }
        if (fault & T2D_FAULT_WRAP_CAP) throw std::runtime_error("seam re-entry did not terminate (EuclideanTiling)");
        double obs[T2D_OBS_LEN];
        check(t2d_observables(gpu, obs), "t2d_observables");
        if (current_step < (int)v_order.size()) v_order[(size_t)current_step] = obs[T2D_OBS_PHI];
        current_step++;
    }
    int32_t slot = -1;
    check(t2d_export_begin(gpu, &slot), "t2d_export_begin");   // returns at once: the copy runs beside the next block
    System landed{};
    if (pending_slot_ >= 0) landed = collect_export(pending_slot_);   // the block before this one, while this one's copy travels
    pending_slot_ = slot;
    if (current_step >= step_count) finished = true;
    return landed;
}

System Tissue2D::flush_export()
{
    System s{};
    if (pending_slot_ >= 0) s = collect_export(pending_slot_);
    pending_slot_ = -1;
    return s;
}

bool Tissue2D::is_finished() { return finished; }

std::vector<double> Tissue2D::get_order_parameter() { return v_order; }

// byte-compatible with IO.h:41-69: append mode, setprecision(15), one row per particle, "," between columns
void Tissue2D::save_our_data()
{
    const size_t N = (size_t)particle_count;
    auto open = [&](const std::string& name) {
        std::ofstream f(ext_.data_dir + "/" + name, std::ios::app);
        if (!f.is_open()) std::cerr << "Error opening file: " << name << std::endl;
        return f;
    };
    {
        std::ofstream f = open("r_data_" + std::to_string(current_step) + ".csv");
        for (size_t i = 0; i < N && f.is_open(); ++i)
            f << std::setprecision(15) << r_UV[i] << "," << std::setprecision(15) << r_UV[N + i] << "\n";
    }
    {
        std::ofstream f = open("r_data_3D_" + std::to_string(current_step) + ".csv");
        for (size_t i = 0; i < N && f.is_open(); ++i)
            f << std::setprecision(15) << r_3D[i] << "," << std::setprecision(15) << r_3D[N + i] << ","
              << std::setprecision(15) << r_3D[2 * N + i] << "\n";
    }
    {
        std::ofstream f = open("particles_color_" + std::to_string(current_step) + ".csv");
        for (size_t i = 0; i < N && f.is_open(); ++i) f << std::setprecision(15) << particles_color[i] << "\n";
    }
}

void Tissue2D::save_state(const std::string& path)
{
    const size_t N = (size_t)particle_count;
    std::vector<double> uv(2 * N);
    std::vector<int32_t> h(N);
    check(t2d_download(gpu, uv.data(), h.data(), nullptr, nullptr, nullptr, nullptr, nullptr), "t2d_download");
    std::ofstream out(path, std::ios::binary);
    if (!out) throw std::runtime_error("cannot write " + path);
    const int64_t hdr[2] = {(int64_t)N, (int64_t)t2d_get_step(gpu)};
    out.write("T2DSTATE", 8);
    out.write(reinterpret_cast<const char*>(hdr), sizeof(hdr));
    out.write(reinterpret_cast<const char*>(uv.data()), sizeof(double) * uv.size());
    out.write(reinterpret_cast<const char*>(h.data()), sizeof(int32_t) * h.size());
}

}  // namespace t2dhost
