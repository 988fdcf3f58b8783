// tissue.h — host-side mirror of the reference's simulation driver `_2DTissue`
// (/root/reference/src/simulation/2DTissue.h:33-60, 2DTissue.cpp:20-280) for the hot path: same constructor
// arguments (order, meaning, defaults), same start() / update() / is_finished() / get_order_parameter() protocol,
// same `Particle` / `System` export (Struct.h:9-29), same exceptions.  perform_particle_simulation() runs on the GPU
// through the C ABI (include/t2d.h); nothing of the step is computed on the host.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/t2d.h"
#include "chart_io.h"

namespace t2dhost {

struct Particle {   // Struct.h:9-22
    double x_UV;
    double y_UV;
    double x_velocity_UV;
    double y_velocity_UV;
    double alignment_UV;
    double compass_north_pole_UV;
    double compass_south_pole_UV;
    double x_3D;
    double y_3D;
    double z_3D;
    int neighbor_count;
};

struct System {     // Struct.h:25-29
    double order_parameter;
    std::vector<Particle> particles;
};

// what the reference hard-wires and this build exposes (SURVEY.md §5 "config / flags")
struct Extensions {
    int neigh_mode = T2D_NEIGH_TABLE;      // --neigh {table,euclid}
    int precision = T2D_PRECISION_FP64;    // --precision {fp64,fp32}
    int lift_mode = T2D_LIFT_REFERENCE;    // --lift {reference,barycentric}
    double eta = 0.0;                      // --noise-eta
    uint64_t seed = 0;                     // --seed (0: std::random_device like CellHelper.cpp:48-49)
    int device = 0;                        // --device
    std::string data_dir = "data";         // where --save-data writes (reference: <PROJECT_SOURCE_DIR>/data)
    std::string load_state;                // binary state to start from instead of init_particle_position
    bool export_particles = true;          // build the Particle vector every step like update() does
    bool quiet = false;                    // suppress the reference's per-step "Step: i" print
    int export_every = 1;                  // --export-every k: the Particle export / CSV of every k-th step only, fetched by
                                           // the asynchronous export (side stream + pinned ring) while the next k steps run
    std::string table_cache;               // --table FILE: a T2DCSR1 binary cache of thresholded table rows (2dtissue_b200/table.py:
                                           // hop counts or the metric geodesic variant) instead of building hop counts from the mesh
    bool device_seed = false;              // --device-seed: t2d_seed_particles (Philox on the GPU) instead of mt19937 on the host
};

class Tissue2D {
  public:
    Tissue2D(bool save_data, bool particle_innenleben, bool free_boundary, std::string mesh_path, int particle_count,
             int step_count = 1, double v0 = 0.1, bool use_kafka = false, double k = 1, double k_next = 10,
             double v0_next = 0.1, double sigma = 0.4166666666666667, double mu = 1, double r_adh = 1, double k_adh = 0.75,
             double step_size = 0.001, int map_cache_count = 30, const Extensions& ext = Extensions());
    ~Tissue2D();
    Tissue2D(const Tissue2D&) = delete;
    Tissue2D& operator=(const Tissue2D&) = delete;

    void start();
    System update();
    // k steps on the device with the previous export's copy still in flight, then the snapshot of this block is started;
    // returns the export that has just landed (the block before), empty for the first block.  run() uses it when export_every > 1
    System update_block(int k);
    System flush_export();   // the last block's export
    bool is_finished();
    std::vector<double> get_order_parameter();

    // checkpoint / resume: uv, heading, step index; the RNG is counter based, so resume is exact
    void save_state(const std::string& path);
    int current_step_index() const { return current_step; }
    const Chart& chart() const { return chart_; }

  private:
    void init_particle_position();   // CellHelper::init_particle_position, CellHelper.cpp:43-67
    void save_our_data();            // _2DTissue::save_our_data, 2DTissue.cpp:270-280
    System collect_export(int slot);
    int pending_slot_ = -1;
    void check(int rc, const char* what);

    bool save_data;
    int particle_count, step_count;
    double v0, k, sigma, step_size;
    int current_step = 0;
    bool finished = false;
    Extensions ext_;
    Chart chart_;
    t2d_ctx* gpu = nullptr;
    // host copies in the reference's Eigen column-major layouts (2DTissue.h:89-106)
    std::vector<double> r_UV, r_dot, r_3D;
    std::vector<int32_t> n, particles_color, vertices_3D_active;
    std::vector<double> v_order;
};

}  // namespace t2dhost
