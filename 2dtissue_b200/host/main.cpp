// main.cpp — `t2d_sim`: the reference's command line (/root/reference/src/main.cpp:20-101) on top of the B200 step.
// Same eight flags with the same defaults and the same quirk (--step-time is wired into v0, main.cpp:50,72,82);
// additional flags expose what the reference hard-wires.  The mesh path names the ORIGINAL mesh like in the
// reference; the chart files its setup wrote next to it (<stem>_uv.off, <stem>_open.off) are what gets loaded.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>

#include "tissue.h"

using namespace t2dhost;

static void usage(const char* argv0)
{
    std::cerr << "Usage: " << argv0 << " [options]\n"
              << "  --step-count N                 Number of steps to simulate (default 20)\n"
              << "  --save-data                    Whether to save simulation data (CSV per step)\n"
              << "  --particle-innenleben          accepted, ignored (as in the reference)\n"
              << "  --optimized-monotile-boundary  accepted, ignored (as in the reference)\n"
              << "  --mesh-path PATH               <stem>.off whose <stem>_uv.off/<stem>_open.off exist, or a .t2dchart\n"
              << "  --particle-count N             Number of particles in the simulation (default 1000)\n"
              << "  --step-time X                  wired into v0 like the reference (default 0.01)\n"
              << "  --kafka                        not available in this build\n"
              << " extensions:\n"
              << "  --neigh {table,euclid}  --precision {fp64,fp32}  --lift {reference,barycentric}  --sigma X  --noise-eta X  --seed N  --device N\n"
              << "  --data-dir DIR  --load-state FILE  --save-state FILE  --dump-chart FILE  --no-particles  --quiet\n"
              << "  --export-every K               export / CSV of every K-th step only, copied asynchronously beside the steps\n"
              << "  --table FILE                   table criterion from a T2DCSR1 cache (hop counts or metric geodesic rows, 2dtissue_b200/table.py)\n"
              << "  --device-seed                  seed the particles on the GPU (Philox) instead of mt19937 on the host\n";
}

int main(int argc, char* argv[])
{
    int step_count = 20, particle_count = 1000;
    bool save_data = false, innenleben = false, monotile = false, use_kafka = false;
    double step_time = 0.01, sigma = 0.4166666666666667;
    std::string mesh_path = "meshes/ellipsoid_x4.off", save_state, dump_chart;
    Extensions ext;
    try {
        for (int i = 1; i < argc; ++i) {
            const std::string a = argv[i];
            auto val = [&]() -> std::string {
                if (i + 1 >= argc) throw std::runtime_error("Too few arguments for '" + a + "'.");
                return argv[++i];
            };
            if (a == "--step-count") step_count = std::stoi(val());
            else if (a == "--save-data") save_data = true;
            else if (a == "--particle-innenleben") innenleben = true;
            else if (a == "--optimized-monotile-boundary") monotile = true;
            else if (a == "--mesh-path") mesh_path = val();
            else if (a == "--particle-count") particle_count = std::stoi(val());
            else if (a == "--step-time") step_time = std::stod(val());
            else if (a == "--kafka") use_kafka = true;
            else if (a == "--neigh") { std::string v = val(); ext.neigh_mode = v == "euclid" ? T2D_NEIGH_EUCLID : T2D_NEIGH_TABLE; if (v != "euclid" && v != "table") throw std::runtime_error("--neigh must be table or euclid"); }
            else if (a == "--precision") { std::string v = val(); ext.precision = v == "fp32" ? T2D_PRECISION_FP32 : T2D_PRECISION_FP64; if (v != "fp32" && v != "fp64") throw std::runtime_error("--precision must be fp64 or fp32"); }
            else if (a == "--lift") { std::string v = val(); ext.lift_mode = v == "barycentric" ? T2D_LIFT_BARYCENTRIC : T2D_LIFT_REFERENCE; if (v != "barycentric" && v != "reference") throw std::runtime_error("--lift must be reference or barycentric"); }
            else if (a == "--sigma") sigma = std::stod(val());
            else if (a == "--noise-eta") ext.eta = std::stod(val());
            else if (a == "--seed") ext.seed = std::stoull(val());
            else if (a == "--device") ext.device = std::stoi(val());
            else if (a == "--data-dir") ext.data_dir = val();
            else if (a == "--load-state") ext.load_state = val();
            else if (a == "--save-state") save_state = val();
            else if (a == "--dump-chart") dump_chart = val();
            else if (a == "--no-particles") ext.export_particles = false;
            else if (a == "--quiet") ext.quiet = true;
            else if (a == "--export-every") ext.export_every = std::max(1, std::stoi(val()));
            else if (a == "--device-seed") ext.device_seed = true;
            else if (a == "--table") ext.table_cache = val();
            else if (a == "-h" || a == "--help") { usage(argv[0]); return 0; }
            else throw std::runtime_error("Unknown argument: " + a);
        }
    } catch (const std::exception& err) {   // argparse: print the error and the usage, EXIT_FAILURE (main.cpp:58-63)
        std::cerr << err.what() << std::endl;
        usage(argv[0]);
        return EXIT_FAILURE;
    }

    try {
        if (!dump_chart.empty()) {   // chart only: no GPU needed
            save_t2dchart(dump_chart, load_chart(mesh_path));
            return 0;
        }
        // v0 = step_time: the reference passes --step-time as the 7th constructor argument (main.cpp:76-84)
        Tissue2D sim(save_data, innenleben, monotile, mesh_path, particle_count, step_count, step_time, use_kafka, 1, 10, 0.1, sigma,
                     1, 1, 0.75, 0.001, 30, ext);
        sim.start();
        const auto t0 = std::chrono::steady_clock::now();
        if (ext.export_every > 1) {   // every-k cadence: asynchronous export, the copy of block b overlaps the steps of block b + 1
            while (!sim.is_finished()) {
                System data = sim.update_block(ext.export_every);
                (void)data;
            }
            System last = sim.flush_export();
            (void)last;
        } else {
            while (!sim.is_finished()) {
                System data = sim.update();
                (void)data;
            }
        }
        for (double v : sim.get_order_parameter()) std::cout << v << '\n';
        const double duration = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::cout << "Time taken: " << duration << " seconds" << '\n';
        if (!save_state.empty()) sim.save_state(save_state);
    } catch (const std::exception& e) {
        std::cerr << "terminate called after throwing an instance of 'std::runtime_error'\n  what():  " << e.what() << std::endl;
        return 134;   // the reference dies with SIGABRT on an uncaught exception
    }
    return 0;
}
