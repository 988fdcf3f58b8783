// api.cu — the extern "C" layer of lib2dtissue_b200.so (include/t2d.h) and the host-side engine behind it.
//
// One context = one GPU = one CUDA stream.  The context owns all device memory: chart (UV triangles, 3-D
// vertices, UV grid), thresholded CSR of the vertex-distance table, cos/sin table, particle SoA (double
// buffered for the counting sort) and the staging areas for the reference-layout host arrays.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <algorithm>
#include <chrono>
#include <functional>
#include <memory>
#include <thread>
#include <string>
#include <vector>

#include "io_kernels.cuh"
#include "t2d_internal.h"

namespace t2d {
int launch_hop_table(int V, const int* d_adj_start, const int* d_adj, uint8_t* out_dev, int sm_count, cudaStream_t s);
int launch_hop_csr(int V, const int* d_adj_start, const int* d_adj, int hops, int* row_len, const int* start, int* col, uint8_t* val,
                   int sm_count, cudaStream_t s);
// comm.cu: NCCL transport (libnccl opened at run time)
struct NcclLink;
int nccl_unique_id(uint8_t* out, std::string* err);
NcclLink* nccl_link_create(int rank, int world, const uint8_t* id_bytes, std::string* err);
void nccl_link_destroy(NcclLink* l);
int nccl_exchange(NcclLink* l, const void* send_left, void* recv_left, const void* send_right, void* recv_right, size_t bytes,
                  const void* far_send, void* far_recv, size_t far_bytes, cudaStream_t s, std::string* err);
}

using namespace t2d;

static std::string g_create_error;

namespace {

struct CudaError {
    std::string msg;
};
#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess) {                                                                            \
            char b__[512];                                                                                   \
            snprintf(b__, sizeof(b__), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            throw CudaError{b__};                                                                            \
        }                                                                                                    \
    } while (0)

template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count)
    {
        release();
        n = count;
        if (count) CK(cudaMalloc((void**)&p, count * sizeof(T)));
    }
    void upload(const std::vector<T>& h, cudaStream_t s)
    {
        alloc(h.size());
        if (!h.empty()) CK(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    }
};

struct EngineBase {
    virtual ~EngineBase() {}
    virtual int set_state(int N, const double* uv, const int* heading, const int* vid, const double* r3d,
                          const uint32_t* ids, bool project) = 0;
    virtual int download(double* uv, int* heading, int* vid, double* r3d, double* rdot, int* color, int* face) = 0;
    virtual int step(int nsteps) = 0;
    virtual int step_host(int N, double* uv, int* heading, int* vid, double* r3d, double* rdot, int* color, bool reproject) = 0;
    virtual int observables(double* out) = 0;
    virtual int get_counters(t2d_counters* out) = 0;
    virtual int reset_counters() = 0;
    virtual int set_tie_log(int on) = 0;
    virtual int seed_particles(int N, uint64_t seed, int mode, uint32_t first_id) = 0;
    virtual int export_begin(int* slot) = 0;
    virtual int export_wait(int slot, int* N, long long* step, const double** uv, const int** heading, const int** vid,
                            const double** r3d, const double** rdot, const int** color) = 0;
    virtual int set_params(const t2d_params* p) = 0;
    virtual int get_r3d(int N, const double* uv, double* r3d, int* vid, int* face) = 0;
    virtual int tiling(int N, double* uv_old, double* uv, int* heading) = 0;
    virtual int unit_vectors(int N, const int* heading, double* out) = 0;
    virtual int forces(double* F, int* new_heading, int* color) = 0;
    virtual int hop_table(uint8_t* out) = 0;
    virtual int profile_step(const char** names, double* ms, int cap) = 0;
    // slab mode
    virtual int comm_init(int rank, int world, const uint8_t* id, const double* cuts, EngineBase* const* group) = 0;
    virtual int comm_destroy() = 0;
    virtual void comm_phase1() = 0;                 // (advance +) classify + pack into the send buffers
    virtual void comm_local_send() = 0;             // local group: copy the messages into the neighbours' receive buffers
    virtual void comm_phase2() = 0;                 // append what arrived + counting sort
    virtual int comm_finish() = 0;                  // synchronise, return the fault mask
    virtual int owned_count() = 0;
    virtual int download_ids(uint32_t* ids) = 0;
    virtual unsigned char* comm_recv_buffer(int dir) = 0;
    virtual unsigned char* comm_far_slot(int src) = 0;   // where rank `src`'s far message lands in this context
    virtual cudaEvent_t comm_event(int which) = 0;   // 0: messages sent, 1: messages consumed
    virtual bool comm_is_local() const = 0;
    virtual bool halo_ready() const = 0;
    virtual int device() const = 0;
    int N = 0;
    int64_t step_index = 0;
    double last_step_ms = 0;
};

// host-side description of the chart + table, precision independent
struct HostChart {
    int V = 0, F = 0;
    std::vector<double> uv, x3d;
    std::vector<int> faces;
    int table_kind = T2D_TABLE_NONE;
    int tableV = 0;
    std::vector<uint8_t> table_raw;   // dense table in the caller's stored type
    // sparse form (T2D_TABLE_CSR_*, or hop counts built on the GPU for large meshes): complete up to csr_radius
    bool sparse = false;
    bool sparse_hops = false;         // rows come from build_hop_csr_device and are rebuilt when the radius grows
    double csr_radius = 0;
    std::vector<int> csr_start, csr_col;
    std::vector<double> csr_val;

    double table_at(int a, int b) const
    {
        size_t k = (size_t)a * tableV + b;
        switch (table_kind) {
            case T2D_TABLE_DENSE_F64: return reinterpret_cast<const double*>(table_raw.data())[k];
            case T2D_TABLE_DENSE_F32: return (double)reinterpret_cast<const float*>(table_raw.data())[k];
            default: return (double)table_raw[k];
        }
    }
};

template <typename R> class Engine : public EngineBase {
  public:
    Engine(const t2d_mesh* mesh, const t2d_table* table, const t2d_params* params, int device);
    ~Engine() override;

    int set_state(int N, const double* uv, const int* heading, const int* vid, const double* r3d, const uint32_t* ids,
                  bool project) override;
    int download(double* uv, int* heading, int* vid, double* r3d, double* rdot, int* color, int* face) override;
    int step(int nsteps) override;
    int step_host(int N, double* uv, int* heading, int* vid, double* r3d, double* rdot, int* color, bool reproject) override;
    int observables(double* out) override;
    int get_counters(t2d_counters* out) override;
    int reset_counters() override;
    int seed_particles(int N, uint64_t seed, int mode, uint32_t first_id) override;
    int export_begin(int* slot) override;
    int export_wait(int slot, int* N, long long* step, const double** uv, const int** heading, const int** vid, const double** r3d,
                    const double** rdot, const int** color) override;
    int set_tie_log(int on) override
    {
        A_.count_ties = on ? 1 : 0;
        return 0;
    }
    int set_params(const t2d_params* p) override;
    int get_r3d(int N, const double* uv, double* r3d, int* vid, int* face) override;
    int tiling(int N, double* uv_old, double* uv, int* heading) override;
    int unit_vectors(int N, const int* heading, double* out) override;
    int forces(double* F, int* new_heading, int* color) override;
    int hop_table(uint8_t* out) override;
    int profile_step(const char** names, double* ms, int cap) override;
    int comm_init(int rank, int world, const uint8_t* id, const double* cuts, EngineBase* const* group) override;
    int comm_destroy() override;
    void comm_phase1() override;
    void comm_local_send() override;
    void comm_phase2() override;
    int comm_finish() override;
    int owned_count() override;
    int download_ids(uint32_t* ids) override;
    unsigned char* comm_recv_buffer(int dir) override { return comm_recv_[dir].p; }
    unsigned char* comm_far_slot(int src) override { return far_recv_.p + (size_t)src * far_bytes_; }
    cudaEvent_t comm_event(int which) override { return which == 0 ? ev_sent_ : ev_consumed_; }
    bool comm_is_local() const override { return comm_on_ && !link_; }
    bool halo_ready() const override { return halo_valid_; }
    int device() const override { return device_; }

  private:
    int compact_owned();   // slab mode: stable compaction offsets of the owned particles; returns their number
    void upload_chart();
    void build_csr();
    void build_vox();
    void alloc_buckets(int nbuckets);
    void build_hop_table_device(std::vector<uint8_t>* host_out);
    void apply_params();
    void resort(bool keys_ready);
    void one_step(bool moving, cudaEvent_t* ev, int* nev);
    int read_fault();
    void ingest(int N, const double* uv, const int* heading, const int* vid, const double* r3d, const uint32_t* ids,
                ParticleArrays<R>& dst);

    int device_ = 0, sm_count_ = 148;
    cudaStream_t stream_ = nullptr;
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    t2d_params P_;
    HostChart chart_;
    int capacity_ = 0;
    bool sorted_ = false;   // `cur` is in bucket order and start[] is valid
    // slab mode
    bool comm_on_ = false, halo_valid_ = false, closing_ = false;
    int resident_ = 0;   // resident slots (owned + halo) at the last compaction
    NcclLink* link_ = nullptr;
    EngineBase* peer_[2] = {nullptr, nullptr};
    std::vector<EngineBase*> group_;   // local slab group: every rank's engine (far channel)
    size_t msg_bytes_ = 0, far_bytes_ = 0;
    DevBuf<unsigned char> comm_send_[2], comm_recv_[2], far_send_, far_recv_;
    DevBuf<DevCommState> comm_state_;
    DevBuf<int> d_cflag_, d_coff_, d_cblk_;
    cudaEvent_t ev_sent_ = nullptr, ev_consumed_ = nullptr;
    std::vector<double> cuts_;
    cudaEvent_t* prof_ev_ = nullptr;   // t2d_profile_step in slab mode: events between the kernels of one step
    int* prof_nev_ = nullptr;
    void prof_mark()
    {
        if (prof_ev_) cudaEventRecord(prof_ev_[(*prof_nev_)++], stream_);
    }
    double vox_sigma_ = -1, vox_color_ = -1;
    double vox_xlo_ = -1e300, vox_xhi_ = 1e300;   // slab mode: only cells that overlap this x range are indexed
    int64_t launches_ = 0, steps_ = 0;

    // chart
    DevBuf<TriUV<R>> d_tri_;
    DevBuf<int4> d_tri_vid_;
    DevBuf<Pos3<R>> d_x3d_;
    DevBuf<int> d_gstart_, d_gfaces_;
    DevBuf<double2> d_trig_d_;
    DevBuf<CrEntry> d_cr_;
    DevBuf<uint2> d_vox_words_;
    DevBuf<int2> d_nbr_;
    DevBuf<float4> d_rec_sentinel_;
    DevBuf<int> d_src_;                          // lean pipeline: sorted slot -> pre-sort index
    DevBuf<unsigned long long> d_scan_status_;   // one-pass scan: tile status words
    unsigned scan_seq_ = 0;                      // launch number of the one-pass scan (tags the status words)
    int scan_ticket_base_ = 0;                   // value of the scan's ticket counter (d_work_[3]) before the next launch
    bool lean_ = false;                          // cur holds only rec + aux (+ src): pos / uv / r_dot / colour are stale until materialize()
    bool lean_ok_ = false;                       // the lean pipeline may be used (fp32 Euclid fast path; T2D_LEAN=0 switches it off)
    void scan_buckets();
    void materialize();
    // asynchronous export: two slots of {device staging, pinned host buffer, "copied" event} and a side stream
    cudaStream_t export_stream_ = nullptr;
    unsigned char* h_export_[2] = {nullptr, nullptr};
    DevBuf<unsigned char> d_export_[2];
    cudaEvent_t ev_export_snap_[2] = {}, ev_export_done_[2] = {};
    int export_n_[2] = {0, 0};
    long long export_step_[2] = {0, 0};
    unsigned export_seq_ = 0;
    // host-buffer path of fp32 contexts: float staging + widening on host threads, chunked so that it overlaps the transfers
    DevBuf<int> d_inv_;             // caller index -> sorted slot (lean sorts write it while step_host32 runs)
    DevBuf<int> d_face_hint_;       // face of every particle (caller order) when the last step_host32 returned
    int face_hint_n_ = -1;          // particle count those hints belong to (-1: none)
    float* h32_ = nullptr;          // pinned: [5N] in (uv, r3d) followed by [7N] out (uv, r3d, rdot)
    size_t h32_cap_ = 0;            // particles the pinned staging holds
    cudaEvent_t ev_chunk_[16] = {};
    int step_host32(int N, double* uv, int* heading, int* vid, double* r3d, double* rdot, int* color, bool reproject);
    void build_hop_csr_device(int hops);
    bool use_fast2_ = false;       // fp32 Euclid: k_step_fast2 (step_fast2.cuh) on 32-byte records; T2D_STEP=legacy switches it off
    DevBuf<int> d_csr_start_, d_csr_col_;
    DevBuf<double> d_csr_d_;
    DevBuf<int> d_adj_start_, d_adj_;
    // particles
    DevBuf<Real2<R>> d_rdot_[2], d_uv_new_, d_F_;
    // what the neighbour search reads of a candidate — pos | cs | uv of one side of the double buffer — lives in ONE
    // allocation, so that an L2 access-policy window can cover it (l2_window)
    DevBuf<unsigned char> d_hot_[2];
    size_t hot_bytes_ = 0;
    size_t l2_persist_bytes_ = 0, l2_window_max_ = 0;
    void l2_window(const void* base);
    DevBuf<int4> d_aux_[2];
    DevBuf<int> d_color_[2], d_new_heading_;
    DevBuf<uint32_t> d_key_, d_rank_;
    DevBuf<int> d_count_, d_start_, d_blocksums_, d_work_;
    DevBuf<DevCounters> d_counters_;
    DevBuf<double> d_obs_;
    DevBuf<unsigned char> d_stage_in_, d_stage_out_;
    StepArgs<R> A_;
};

// ---------------------------------------------------------------------------------------------------
template <typename R> Engine<R>::Engine(const t2d_mesh* mesh, const t2d_table* table, const t2d_params* params, int device)
{
    device_ = device;
    P_ = *params;
    CK(cudaSetDevice(device_));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device_));
    if (prop.major < 10) throw CudaError{"lib2dtissue_b200 is built for sm_100a (B200) only; device is sm_" +
                                         std::to_string(prop.major) + std::to_string(prop.minor)};
    sm_count_ = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    CK(cudaEventCreate(&ev0_));
    CK(cudaEventCreate(&ev1_));

    if (!mesh || mesh->V <= 0 || mesh->F <= 0 || !mesh->uv || !mesh->x3d || !mesh->faces) throw CudaError{"invalid mesh"};
    chart_.V = mesh->V;
    chart_.F = mesh->F;
    chart_.uv.assign(mesh->uv, mesh->uv + 2 * (size_t)mesh->V);
    chart_.x3d.assign(mesh->x3d, mesh->x3d + 3 * (size_t)mesh->V);
    chart_.faces.assign(mesh->faces, mesh->faces + 3 * (size_t)mesh->F);
    for (int v : chart_.faces)
        if (v < 0 || v >= chart_.V) throw CudaError{"face index out of range"};

    if (table && table->kind != T2D_TABLE_NONE) {
        chart_.table_kind = table->kind;
        chart_.tableV = table->V;
        if (table->V != mesh->V) throw CudaError{"table size does not match the mesh's vertex count"};
        size_t nn = (size_t)table->V * table->V;
        if (table->kind == T2D_TABLE_HOPS_FROM_MESH) {
            // filled after the chart upload (needs the adjacency on the device)
        } else if (table->kind == T2D_TABLE_CSR_F64 || table->kind == T2D_TABLE_CSR_F32 || table->kind == T2D_TABLE_CSR_U8) {
            const t2d_table_csr* c = static_cast<const t2d_table_csr*>(table->data);
            if (!c || !c->start || !c->col || !c->val || c->nnz < 0) throw CudaError{"table CSR is null"};
            const int V = table->V;
            if (c->start[0] != 0 || (int64_t)c->start[V] != c->nnz) throw CudaError{"table CSR: start[] does not match nnz"};
            chart_.sparse = true;
            chart_.csr_radius = c->radius;
            chart_.csr_start.assign(c->start, c->start + V + 1);
            chart_.csr_col.assign(c->col, c->col + c->nnz);
            chart_.csr_val.resize((size_t)c->nnz);
            for (int64_t q = 0; q < c->nnz; ++q)
                chart_.csr_val[(size_t)q] = table->kind == T2D_TABLE_CSR_F64 ? static_cast<const double*>(c->val)[q]
                                            : table->kind == T2D_TABLE_CSR_F32 ? (double)static_cast<const float*>(c->val)[q]
                                                                               : (double)static_cast<const uint8_t*>(c->val)[q];
            for (int v = 0; v < V; ++v) {
                if (c->start[v + 1] < c->start[v]) throw CudaError{"table CSR: start[] must be non-decreasing"};
                bool diag = false;
                for (int q = c->start[v]; q < c->start[v + 1]; ++q) {
                    const int u = c->col[q];
                    if (u < 0 || u >= V || (q > c->start[v] && u <= c->col[q - 1])) throw CudaError{"table CSR: columns must ascend inside a row"};
                    if (u == v) {
                        diag = true;
                        if (chart_.csr_val[(size_t)q] != 0.0) throw CudaError{"vertex-distance table must have a zero diagonal"};
                    }
                }
                if (!diag) throw CudaError{"table CSR: every row must hold its diagonal entry"};
            }
        } else {
            if (!table->data) throw CudaError{"table data is null"};
            size_t es = table->kind == T2D_TABLE_DENSE_F64 ? 8 : (table->kind == T2D_TABLE_DENSE_F32 ? 4 : 1);
            const uint8_t* src = static_cast<const uint8_t*>(table->data);
            chart_.table_raw.assign(src, src + nn * es);
        }
    }
    if (P_.neigh_mode == T2D_NEIGH_TABLE && chart_.table_kind == T2D_TABLE_NONE)
        throw CudaError{"neigh_mode = table needs a vertex-distance table"};

    capacity_ = P_.capacity > 0 ? P_.capacity : 1;
    {
        const char* e = getenv("T2D_STEP");   // dev knob for A/B measurements: "legacy" = k_step_euclid_fast + the cs array
        use_fast2_ = sizeof(R) == 4 && P_.neigh_mode == T2D_NEIGH_EUCLID && !(e && std::string(e) == "legacy");
        const char* t = getenv("T2D_COUNT_TIES");
        A_.count_ties = (t && atoi(t) != 0) ? 1 : 0;   // the tie log is a diagnostic: off unless asked for (t2d_set_tie_log)
    }
    upload_chart();
    if (chart_.table_kind == T2D_TABLE_HOPS_FROM_MESH) {
        const char* fs = getenv("T2D_HOPS_SPARSE");   // dev/test knob: force the sparse builder on a small mesh
        if ((double)chart_.V * chart_.V > 1.0e9 || (fs && atoi(fs) != 0)) {
            chart_.sparse = chart_.sparse_hops = true;   // rows are built (and rebuilt) by apply_params for the radius in force
            chart_.tableV = chart_.V;
        } else {
            build_hop_table_device(&chart_.table_raw);
            chart_.table_kind = T2D_TABLE_DENSE_U8;
        }
    }

    // particle storage
    const size_t C = (size_t)capacity_;
    for (int b = 0; b < 2; ++b) {
        const bool with_cs = sizeof(R) == 4 && P_.neigh_mode == T2D_NEIGH_EUCLID && !use_fast2_;
        hot_bytes_ = C * (sizeof(Pos3<R>) + (with_cs ? sizeof(double2) : 0) + (use_fast2_ ? 2 * sizeof(float4) : 0) + sizeof(Real2<R>)) + 64;
        d_hot_[b].alloc(hot_bytes_);
        d_aux_[b].alloc(C);
        d_rdot_[b].alloc(C);
        d_color_[b].alloc(C);
    }
    d_uv_new_.alloc(C);
    d_F_.alloc(C);
    d_new_heading_.alloc(C);
    d_key_.alloc(C);
    d_rank_.alloc(C);
    d_counters_.alloc(1);
    CK(cudaMemsetAsync(d_counters_.p, 0, sizeof(DevCounters), stream_));
    d_obs_.alloc(T2D_OBS_LEN);
    d_work_.alloc(4);
    CK(cudaMemsetAsync(d_work_.p, 0, 4 * sizeof(int), stream_));
    if (use_fast2_) {   // {x, y, z, trig slot 361 = zero entry | u, v, cell, heading}: d^2 overflows to +inf against any real particle
        int slot = 361;
        float sl;
        memcpy(&sl, &slot, 4);
        const float4 h[2] = {make_float4(1e30f, 1e30f, 1e30f, sl), make_float4(0.f, 0.f, 0.f, 0.f)};
        d_rec_sentinel_.alloc(2);
        CK(cudaMemcpyAsync(d_rec_sentinel_.p, h, sizeof(h), cudaMemcpyHostToDevice, stream_));
        CK(cudaStreamSynchronize(stream_));
        A_.rec_sentinel = d_rec_sentinel_.p;
        d_src_.alloc(C);
        A_.src = d_src_.p;
        if (const char* pe = getenv("T2D_PDL")) A_.pdl = atoi(pe) != 0;
        const char* le = getenv("T2D_LEAN");
        lean_ok_ = !(le && atoi(le) == 0);
    }
    d_stage_in_.alloc(C * (16 + 4 + 4 + 24 + 4) + 256);
    d_stage_out_.alloc(C * (16 + 4 + 4 + 24 + 16 + 4 + 4 + 16 + 4) + 256);

    auto carve = [&](int b) {
        const bool with_cs = sizeof(R) == 4 && P_.neigh_mode == T2D_NEIGH_EUCLID && !use_fast2_;
        unsigned char* q = d_hot_[b].p;
        Pos3<R>* pos = reinterpret_cast<Pos3<R>*>(q);
        q += C * sizeof(Pos3<R>);
        double2* cs = with_cs ? reinterpret_cast<double2*>(q) : nullptr;
        if (with_cs) q += C * sizeof(double2);
        float4* rec = use_fast2_ ? reinterpret_cast<float4*>(q) : nullptr;
        if (use_fast2_) q += C * 2 * sizeof(float4);
        Real2<R>* uv = reinterpret_cast<Real2<R>*>(q);
        return ParticleArrays<R>{pos, uv, d_aux_[b].p, d_rdot_[b].p, d_color_[b].p, cs, rec};
    };
    A_.cur = carve(0);
    A_.alt = carve(1);
    {   // optional L2 set-aside for the sorted candidate data: T2D_L2_PERSIST=1 (or a size in MB).  Off by default —
        // measured on the bench workload: 0.472 ms per step without, 0.479-0.493 with (32 MB / 64 MB / maximum)
        const char* e = getenv("T2D_L2_PERSIST");
        if (e && atoi(e) >= 1 && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
            l2_persist_bytes_ = (size_t)prop.persistingL2CacheMaxSize;
            if (e && atoi(e) > 1) l2_persist_bytes_ = std::min(l2_persist_bytes_, (size_t)atoi(e) << 20);   // MB (dev knob)
            l2_window_max_ = (size_t)prop.accessPolicyMaxWindowSize;
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, l2_persist_bytes_) != cudaSuccess) {
                cudaGetLastError();
                l2_persist_bytes_ = 0;
            }
        }
    }
    A_.key = d_key_.p;
    A_.rank = d_rank_.p;
    A_.uv_new = d_uv_new_.p;
    A_.F = d_F_.p;
    A_.new_heading = d_new_heading_.p;
    A_.counters = d_counters_.p;
    A_.trig_d = d_trig_d_.p;
    A_.cr = d_cr_.p;
    A_.work_counter = d_work_.p;
    A_.mode = P_.neigh_mode;
    A_.mesh.lift_mode = P_.lift_mode;
    A_.write_F = 0;
    apply_params();
    CK(cudaStreamSynchronize(stream_));
}

template <typename R> Engine<R>::~Engine()
{
    if (h32_) cudaFreeHost(h32_);
    for (int k = 0; k < 2; ++k) {
        if (h_export_[k]) cudaFreeHost(h_export_[k]);
        if (ev_export_snap_[k]) cudaEventDestroy(ev_export_snap_[k]);
        if (ev_export_done_[k]) cudaEventDestroy(ev_export_done_[k]);
    }
    if (export_stream_) cudaStreamDestroy(export_stream_);
    for (auto& e : ev_chunk_)
        if (e) cudaEventDestroy(e);
    closing_ = true;
    comm_destroy();
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
    if (stream_) cudaStreamDestroy(stream_);
}

// chart -> device: packed UV triangles, 3-D vertices, uniform UV grid (cell -> faces whose grown bounding
// box touches the cell, ascending), cos/sin table from the host's libm, vertex adjacency for the BFS
template <typename R> void Engine<R>::upload_chart()
{
    const int V = chart_.V, F = chart_.F;
    std::vector<TriUV<R>> tri(F);
    std::vector<int4> tv(F);
    for (int f = 0; f < F; ++f) {
        const int* fv = &chart_.faces[3 * (size_t)f];
        tri[f].ax = (R)chart_.uv[2 * (size_t)fv[0]];
        tri[f].ay = (R)chart_.uv[2 * (size_t)fv[0] + 1];
        tri[f].bx = (R)chart_.uv[2 * (size_t)fv[1]];
        tri[f].by = (R)chart_.uv[2 * (size_t)fv[1] + 1];
        tri[f].cx = (R)chart_.uv[2 * (size_t)fv[2]];
        tri[f].cy = (R)chart_.uv[2 * (size_t)fv[2] + 1];
        tv[f] = make_int4(fv[0], fv[1], fv[2], 0);
    }
    std::vector<Pos3<R>> x3(V);
    double mn[3] = {1e300, 1e300, 1e300};
    for (int v = 0; v < V; ++v) {
        x3[v].x = (R)chart_.x3d[3 * (size_t)v];
        x3[v].y = (R)chart_.x3d[3 * (size_t)v + 1];
        x3[v].z = (R)chart_.x3d[3 * (size_t)v + 2];
        x3[v].w = R(0);
        for (int k = 0; k < 3; ++k) mn[k] = std::min(mn[k], chart_.x3d[3 * (size_t)v + k]);
    }
    d_tri_.upload(tri, stream_);
    d_tri_vid_.upload(tv, stream_);
    d_x3d_.upload(x3, stream_);

    int G = (int)ceil(sqrt((double)F) * 2.0);
    G = std::max(8, std::min(G, 4096));
    const double eps = 1e-6;
    std::vector<int> cnt((size_t)G * G + 1, 0);
    auto cell_range = [&](int f, int& i0, int& i1, int& j0, int& j1) {
        const int* fv = &chart_.faces[3 * (size_t)f];
        double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
        for (int k = 0; k < 3; ++k) {
            double x = chart_.uv[2 * (size_t)fv[k]], y = chart_.uv[2 * (size_t)fv[k] + 1];
            x0 = std::min(x0, x);
            x1 = std::max(x1, x);
            y0 = std::min(y0, y);
            y1 = std::max(y1, y);
        }
        i0 = std::max(0, (int)floor((x0 - eps) * G));
        i1 = std::min(G - 1, (int)floor((x1 + eps) * G));
        j0 = std::max(0, (int)floor((y0 - eps) * G));
        j1 = std::min(G - 1, (int)floor((y1 + eps) * G));
    };
    for (int f = 0; f < F; ++f) {
        int i0, i1, j0, j1;
        cell_range(f, i0, i1, j0, j1);
        for (int j = j0; j <= j1; ++j)
            for (int i = i0; i <= i1; ++i) cnt[(size_t)j * G + i + 1]++;
    }
    std::vector<int> gstart((size_t)G * G + 1, 0);
    for (size_t c = 0; c < (size_t)G * G; ++c) gstart[c + 1] = gstart[c] + cnt[c + 1];
    std::vector<int> gfaces((size_t)std::max(1, gstart[(size_t)G * G]));
    std::fill(cnt.begin(), cnt.end(), 0);
    for (int f = 0; f < F; ++f) {   // ascending f => every cell list is ascending
        int i0, i1, j0, j1;
        cell_range(f, i0, i1, j0, j1);
        for (int j = j0; j <= j1; ++j)
            for (int i = i0; i <= i1; ++i) {
                size_t c = (size_t)j * G + i;
                gfaces[(size_t)gstart[c] + cnt[c]++] = f;
            }
    }
    d_gstart_.upload(gstart, stream_);
    d_gfaces_.upload(gfaces, stream_);
    A_.mesh.V = V;
    A_.mesh.F = F;
    A_.mesh.G = G;
    A_.mesh.tri = d_tri_.p;
    A_.mesh.tri_vid = d_tri_vid_.p;
    A_.mesh.x3d = d_x3d_.p;
    A_.mesh.gstart = d_gstart_.p;
    A_.mesh.gfaces = d_gfaces_.p;

    // cos/sin of integer degrees exactly as the reference's libm call sees them:
    // cos(double(n) * DEG_TO_RAD), LinearAlgebra.cpp:37-41, OrientationHelper.cpp:96-100
    std::vector<double2> td(TRIG_N);
    for (int q = 0; q < TRIG_N; ++q) {
        double angle_degrees = (double)(TRIG_MIN + q);
        double angle_radians = angle_degrees * DEG_TO_RAD_D;
        td[q] = make_double2(cos(angle_radians), sin(angle_radians));
    }
    d_trig_d_.upload(td, stream_);
    std::vector<CrEntry> cr(kCrTable, kCrTable + 181);
    d_cr_.upload(cr, stream_);

    // vertex adjacency (both directions of every face edge; duplicates are harmless for BFS)
    std::vector<int> deg((size_t)V + 1, 0);
    for (int f = 0; f < F; ++f)
        for (int k = 0; k < 3; ++k) deg[(size_t)chart_.faces[3 * (size_t)f + k] + 1] += 2;
    std::vector<int> as((size_t)V + 1, 0);
    for (int v = 0; v < V; ++v) as[v + 1] = as[v] + deg[v + 1];
    std::vector<int> adj((size_t)std::max(1, as[V])), fill((size_t)V, 0);
    for (int f = 0; f < F; ++f)
        for (int k = 0; k < 3; ++k) {
            int a = chart_.faces[3 * (size_t)f + k], b = chart_.faces[3 * (size_t)f + (k + 1) % 3];
            adj[(size_t)as[a] + fill[a]++] = b;
            adj[(size_t)as[b] + fill[b]++] = a;
        }
    d_adj_start_.upload(as, stream_);
    d_adj_.upload(adj, stream_);
    CK(cudaStreamSynchronize(stream_));
}

template <typename R> void Engine<R>::build_hop_table_device(std::vector<uint8_t>* host_out)
{
    const int V = chart_.V;
    DevBuf<uint8_t> d_tab;
    d_tab.alloc((size_t)V * V);
    int e = launch_hop_table(V, d_adj_start_.p, d_adj_.p, d_tab.p, sm_count_, stream_);
    if (e != 0) throw CudaError{std::string("hop-table kernel launch failed: ") + cudaGetErrorString((cudaError_t)e)};
    launches_++;
    host_out->resize((size_t)V * V);
    CK(cudaMemcpyAsync(host_out->data(), d_tab.p, (size_t)V * V, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    chart_.tableV = V;
}

// hop counts up to `hops` as CSR rows, built on the GPU without ever forming the V x V table (hop_table.cu): count pass,
// host scan of the V row lengths, fill pass
template <typename R> void Engine<R>::build_hop_csr_device(int hops)
{
    const int V = chart_.V;
    hops = std::max(1, std::min(hops, 254));
    DevBuf<int> d_len, d_start, d_col;
    DevBuf<uint8_t> d_val;
    d_len.alloc((size_t)V + 1);
    int e = launch_hop_csr(V, d_adj_start_.p, d_adj_.p, hops, d_len.p, nullptr, nullptr, nullptr, sm_count_, stream_);
    if (e != 0) throw CudaError{std::string("hop-CSR kernel launch failed: ") + cudaGetErrorString((cudaError_t)e)};
    std::vector<int> len((size_t)V);
    CK(cudaMemcpyAsync(len.data(), d_len.p, (size_t)V * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    chart_.csr_start.assign((size_t)V + 1, 0);
    long long total = 0;
    for (int v = 0; v < V; ++v) {
        total += len[(size_t)v];
        if (total > 2000000000LL) throw CudaError{"hop CSR too large: lower sigma or use a coarser chart"};
        chart_.csr_start[(size_t)v + 1] = (int)total;
    }
    d_start.upload(chart_.csr_start, stream_);
    d_col.alloc((size_t)total + 1);
    d_val.alloc((size_t)total + 1);
    e = launch_hop_csr(V, d_adj_start_.p, d_adj_.p, hops, nullptr, d_start.p, d_col.p, d_val.p, sm_count_, stream_);
    if (e != 0) throw CudaError{std::string("hop-CSR kernel launch failed: ") + cudaGetErrorString((cudaError_t)e)};
    launches_ += 2;
    chart_.csr_col.resize((size_t)total);
    std::vector<uint8_t> val((size_t)total);
    CK(cudaMemcpyAsync(chart_.csr_col.data(), d_col.p, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(val.data(), d_val.p, (size_t)total, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    chart_.csr_val.assign(val.begin(), val.end());
    chart_.csr_radius = (double)hops;
}

// thresholded CSR of the table: entries with d < 2σ or d <= color_factor·σ, d = min(D[v][u], D[u][v])
// (Locomotion.cpp:110 symmetrises by min), kept as doubles — the reference's in-memory precision.
template <typename R> void Engine<R>::build_csr()
{
    const int V = chart_.tableV;
    const double two_sigma = 2 * P_.sigma, color_r = P_.color_factor * P_.sigma;
    std::vector<int> start((size_t)V + 1, 0), col;
    std::vector<double> dv;
    if (chart_.sparse) {   // the caller's (or the GPU builder's) rows, cut to the radius in force
        const double rmax = std::max(two_sigma, color_r);
        if (chart_.sparse_hops && (chart_.csr_start.empty() || rmax > chart_.csr_radius)) build_hop_csr_device((int)std::ceil(rmax));
        if (rmax > chart_.csr_radius)
            throw CudaError{"the table CSR is complete up to its radius only: max(2 sigma, color_factor sigma) exceeds it"};
        for (int v = 0; v < V; ++v) {
            for (int q = chart_.csr_start[v]; q < chart_.csr_start[v + 1]; ++q) {
                const double d = chart_.csr_val[(size_t)q];
                if (d < two_sigma || (d != 0.0 && d <= color_r)) {
                    col.push_back(chart_.csr_col[(size_t)q]);
                    dv.push_back(d);
                }
            }
            start[v + 1] = (int)col.size();
        }
    } else
    for (int v = 0; v < V; ++v) {
        if (chart_.table_at(v, v) != 0.0) throw CudaError{"vertex-distance table must have a zero diagonal"};
        for (int u = 0; u < V; ++u) {
            double d = chart_.table_at(v, u), dT = chart_.table_at(u, v);
            if (dT < d) d = dT;
            if (d < two_sigma || (d != 0.0 && d <= color_r)) {
                col.push_back(u);
                dv.push_back(d);
            }
        }
        start[v + 1] = (int)col.size();
    }
    if (col.empty()) {
        col.push_back(0);
        dv.push_back(0.0);
    }
    d_csr_start_.upload(start, stream_);
    d_csr_col_.upload(col, stream_);
    d_csr_d_.upload(dv, stream_);
    CK(cudaStreamSynchronize(stream_));
    A_.csr.V = V;
    A_.csr.start = d_csr_start_.p;
    A_.csr.col = d_csr_col_.p;
    A_.csr.d = d_csr_d_.p;
}

template <typename R> void Engine<R>::alloc_buckets(int nbuckets)
{
    A_.M = nbuckets;
    if (d_count_.n < (size_t)nbuckets + 8) {
        d_count_.alloc((size_t)nbuckets + 8);
        d_start_.alloc((size_t)nbuckets + 8);
        d_blocksums_.alloc((size_t)scan_blocks(nbuckets) + 8);
    }
    CK(cudaMemsetAsync(d_count_.p, 0, d_count_.n * sizeof(int), stream_));
    A_.count = d_count_.p;
    A_.start = d_start_.p;
    A_.blocksums = d_blocksums_.p;
    d_scan_status_.alloc((size_t)scan_blocks(nbuckets) + 8);
    CK(cudaMemsetAsync(d_scan_status_.p, 0, ((size_t)scan_blocks(nbuckets) + 8) * sizeof(unsigned long long), stream_));
}

// exclusive scan of the bucket histogram into A_.start (zeroes the histogram): one launch
template <typename R> void Engine<R>::scan_buckets()
{
    const int nb = scan_blocks(A_.M);
    if (scan_ticket_base_ > (1 << 30)) {   // keep the ticket counter far from overflow
        CK(cudaMemsetAsync(d_work_.p + 3, 0, sizeof(int), stream_));
        scan_ticket_base_ = 0;
    }
    launch_scan_onepass(A_.count, A_.start, d_scan_status_.p, d_work_.p + 3, scan_ticket_base_, ++scan_seq_, A_.M, stream_, A_.pdl != 0);
    scan_ticket_base_ += nb;
    launches_++;
}

// lean pipeline -> full sorted state (pos, uv, r_dot, colour), for everything that is not k_step_fast2
template <typename R> void Engine<R>::materialize()
{
    if (!lean_) return;
    Launch<R>::expand(A_, stream_);
    launches_++;
    lean_ = false;
}

// Sparse row index of the 3-D cell list (t2d_internal.h DevVox).  Cell edge = rmax*(1+margin); the cells a
// particle can ever occupy are those the mesh surface touches (positions are convex combinations of a face's
// corners, CellHelper.cpp:143-146), found by a GPU voxelisation.  Compact indices ascend along x inside a row; the
// rows follow a Morton curve over (y, z) (T2D_ROW_ORDER=lex: plain (z, y) order, for A/B measurements).
template <typename R> void Engine<R>::build_vox()
{
    const double two_sigma = 2 * P_.sigma, color_r = P_.color_factor * P_.sigma;
    const double rmax = std::max(two_sigma, color_r);
    const double margin = sizeof(R) == 8 ? 1e-9 : 1e-3;
    const double cs = rmax * (1.0 + margin);
    if (!(cs > 0)) throw CudaError{"sigma must be positive"};
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int v = 0; v < chart_.V; ++v)
        for (int k = 0; k < 3; ++k) {
            mn[k] = std::min(mn[k], chart_.x3d[3 * (size_t)v + k]);
            mx[k] = std::max(mx[k], chart_.x3d[3 * (size_t)v + k]);
        }
    double org[3];
    int nc[3];
    double extent = 0;
    for (int k = 0; k < 3; ++k) {
        org[k] = mn[k] - 2.0 * cs;
        double n = std::floor((mx[k] - org[k]) / cs) + 3.0;
        if (n > 2.0e9) throw CudaError{"sigma is too small for the mesh extent (cell grid axis overflows)"};
        nc[k] = (int)n;
        extent = std::max(extent, mx[k] - mn[k]);
    }
    const int nwx = (nc[0] >> 5) + 1;   // + a spare word at the end of every row (row_range reads rank(ncx))
    const size_t nrows = (size_t)nc[1] * nc[2];
    const double nwords_d = (double)nrows * nwx;
    if (nwords_d > 7.5e8) throw CudaError{"sigma is too small for the mesh extent (the row table of the cell list would exceed 6 GB)"};
    const size_t nwords = (size_t)nwords_d;
    const double reach = 0.5 * std::sqrt(3.0) * cs + 0.02 * cs + 2e-5 * extent;

    DevBuf<unsigned> d_occ;
    d_occ.alloc(nwords);
    CK(cudaMemsetAsync(d_occ.p, 0, nwords * sizeof(unsigned), stream_));
    Launch<R>::voxelize(A_.mesh, org, cs, reach, nc, nwx, d_occ.p, stream_);
    launches_++;
    std::vector<unsigned> occ(nwords);
    CK(cudaMemcpyAsync(occ.data(), d_occ.p, nwords * sizeof(unsigned), cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    CK(cudaGetLastError());

    // slab mode: a rank only ever holds particles of its slab and its halo strips, so cells outside that x range are
    // dropped from the index — the bucket arrays (histogram, scan) then scale with the rank's particles, not the job's
    if (vox_xlo_ > -1e299 || vox_xhi_ < 1e299) {
        const int cx_lo = (int)std::max(0.0, std::floor((vox_xlo_ - org[0]) / cs)), cx_hi = (int)std::min((double)nc[0] - 1, std::floor((vox_xhi_ - org[0]) / cs));
        for (size_t r = 0; r < nrows; ++r)
            for (int w = 0; w < nwx; ++w) {
                unsigned& bits = occ[r * nwx + w];
                if (!bits) continue;
                const int c0 = w * 32;
                if (c0 + 31 < cx_lo || c0 > cx_hi) {
                    bits = 0;
                } else {
                    unsigned keep = 0xffffffffu;
                    if (cx_lo > c0) keep &= 0xffffffffu << (cx_lo - c0);
                    if (cx_hi < c0 + 31) keep &= 0xffffffffu >> (c0 + 31 - cx_hi);
                    bits &= keep;
                }
            }
    }

    // non-empty rows in Morton order over (y, z) -> compact base index of every word
    auto spread = [](uint64_t v) {   // 32 bits -> every second bit
        v &= 0xffffffffull;
        v = (v | v << 16) & 0x0000ffff0000ffffull;
        v = (v | v << 8) & 0x00ff00ff00ff00ffull;
        v = (v | v << 4) & 0x0f0f0f0f0f0f0f0full;
        v = (v | v << 2) & 0x3333333333333333ull;
        v = (v | v << 1) & 0x5555555555555555ull;
        return v;
    };
    const char* ord = getenv("T2D_ROW_ORDER");
    const bool lex = ord && std::string(ord) == "lex";
    std::vector<std::pair<uint64_t, size_t>> order;
    for (size_t r = 0; r < nrows; ++r) {
        bool any = false;
        for (int w = 0; w < nwx && !any; ++w) any = occ[r * nwx + w] != 0;
        if (!any) continue;
        const uint64_t y = r % nc[1], z = r / nc[1];
        order.emplace_back(lex ? (uint64_t)r : (spread(y) | (spread(z) << 1)), r);
    }
    std::sort(order.begin(), order.end());
    std::vector<uint2> words(nwords, make_uint2(0, 0));
    long long base = 0;
    for (auto& o : order)
        for (int w = 0; w < nwx; ++w) {
            const unsigned bits = occ[o.second * nwx + w];
            words[o.second * nwx + w] = make_uint2(bits, (unsigned)base);
            base += __builtin_popcount(bits);
        }
    if (base > 2000000000LL) throw CudaError{"cell list too large"};
    d_vox_words_.upload(words, stream_);
    CK(cudaStreamSynchronize(stream_));
    A_.vox.ncx = nc[0];
    A_.vox.ncy = nc[1];
    A_.vox.ncz = nc[2];
    A_.vox.nwx = nwx;
    A_.vox.M = (int)base;
    A_.vox.words = d_vox_words_.p;
    for (int k = 0; k < 3; ++k) A_.vox.origin[k] = (R)org[k];
    A_.vox.inv_cell = (R)(1.0 / cs);
    alloc_buckets((int)base + 1);   // + the overflow bucket
    if (use_fast2_) {   // static neighbourhood table of every compact cell (+ the overflow bucket's unused entry)
        d_nbr_.alloc(((size_t)base + 1) * NBR_STRIDE);
        Launch<R>::build_nbr(A_.vox, d_nbr_.p, stream_);
        launches_++;
        CK(cudaStreamSynchronize(stream_));
        CK(cudaGetLastError());
        A_.nbr = d_nbr_.p;
    }
}

template <typename R> void Engine<R>::apply_params()
{
    A_.v0 = (R)P_.v0;
    A_.k = (R)P_.k;
    const double two_sigma = 2 * P_.sigma;   // "2 * σ" in double, ForceHelper.cpp:55, OrientationHelper.cpp:58
    const double color_r = P_.color_factor * P_.sigma;   // "2.4 * σ", 2DTissue.cpp:262
    A_.two_sigma = (R)two_sigma;
    A_.color_r = (R)color_r;
    A_.two_sigma_d = two_sigma;
    A_.color_r_d = color_r;
    A_.step_size = (R)P_.step_size;
    A_.eta360 = P_.eta * 360.0;
    A_.seed = P_.seed;
    if (P_.neigh_mode == T2D_NEIGH_EUCLID) {
        if (vox_sigma_ != P_.sigma || vox_color_ != P_.color_factor) {
            build_vox();
            vox_sigma_ = P_.sigma;
            vox_color_ = P_.color_factor;
            sorted_ = false;
        }
    } else {
        build_csr();
        alloc_buckets(chart_.V);
        sorted_ = false;
    }
}

template <typename R> int Engine<R>::set_params(const t2d_params* p)
{
    if (p->neigh_mode != P_.neigh_mode || p->precision != P_.precision || p->lift_mode != P_.lift_mode)
        throw CudaError{"neigh_mode/precision/lift_mode are fixed at create"};
    // slab mode: the halo width, the 4 r_max slab-width check and the x range of the cell index were all derived from
    // r_max at t2d_comm_init; a new interaction radius would silently miss neighbours across a cut
    if (comm_on_ && (p->sigma != P_.sigma || p->color_factor != P_.color_factor))
        throw CudaError{"sigma / color_factor cannot change while the slab exchange is active: t2d_comm_destroy, "
                        "t2d_set_params, then t2d_comm_init again"};
    CK(cudaSetDevice(device_));
    materialize();
    int cap = P_.capacity;
    P_ = *p;
    P_.capacity = cap;
    apply_params();
    return 0;
}

template <typename R>
void Engine<R>::ingest(int N, const double* uv, const int* heading, const int* vid, const double* r3d, const uint32_t* ids,
                       ParticleArrays<R>& dst)
{
    // raw host arrays -> device staging (sizes of the reference's own arrays) -> SoA conversion on the device
    unsigned char* base = d_stage_in_.p;
    size_t off = 0;
    HostViewIn in{};
    auto put = [&](const void* src, size_t bytes) -> void* {
        void* d = base + off;
        CK(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, stream_));
        off += (bytes + 15) & ~(size_t)15;
        return d;
    };
    in.uv = (const double*)put(uv, sizeof(double) * 2 * (size_t)N);
    in.heading = heading ? (const int*)put(heading, sizeof(int) * (size_t)N) : nullptr;
    in.vid = vid ? (const int*)put(vid, sizeof(int) * (size_t)N) : nullptr;
    in.r3d = r3d ? (const double*)put(r3d, sizeof(double) * 3 * (size_t)N) : nullptr;
    in.ids = ids ? (const uint32_t*)put(ids, sizeof(uint32_t) * (size_t)N) : nullptr;
    IoLaunch<R>::ingest(N, in, dst, stream_);
    launches_++;
}

// The counting-sort scatter writes the sorted pos | cs | uv block that the NEXT step kernel gathers its candidates from.
// An access-policy window on that block keeps as much of it as the L2 set-aside holds from being evicted by the
// streaming traffic in between (the unsorted new state, aux, rdot); everything outside the window is "streaming".
template <typename R> void Engine<R>::l2_window(const void* base)
{
    if (!l2_persist_bytes_ || this->N <= 0) return;
    const size_t used = std::min(hot_bytes_, (size_t)(comm_on_ ? capacity_ : this->N) * (hot_bytes_ / (size_t)capacity_) + 64);
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    v.accessPolicyWindow.base_ptr = const_cast<void*>(base);
    v.accessPolicyWindow.num_bytes = std::min(used, l2_window_max_);
    v.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)l2_persist_bytes_ / (double)v.accessPolicyWindow.num_bytes);
    v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(stream_, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) {
        cudaGetLastError();
        l2_persist_bytes_ = 0;   // not supported here: never try again
    }
}

// counting sort of `cur` into bucket order (cur -> alt, then swap); keys_ready: the producing kernel already
// wrote key / rank / histogram
template <typename R> void Engine<R>::resort(bool keys_ready)
{
    if (!keys_ready) {
        Launch<R>::bin(A_, stream_);
        launches_++;
    }
    scan_buckets();
    l2_window(A_.alt.pos);
    Launch<R>::scatter(A_, stream_);
    std::swap(A_.cur, A_.alt);
    launches_ += 1;
    sorted_ = true;
    lean_ = false;
}

template <typename R>
int Engine<R>::set_state(int N, const double* uv, const int* heading, const int* vid, const double* r3d, const uint32_t* ids,
                         bool project)
{
    if (N < 0 || N > capacity_) throw CudaError{"particle count exceeds the context's capacity"};
    if (!uv || !heading) throw CudaError{"uv and heading are required"};
    if (vid && P_.neigh_mode == T2D_NEIGH_TABLE)
        for (int i = 0; i < N; ++i)
            if (vid[i] < 0 || vid[i] >= chart_.V) throw CudaError{"vid out of range"};
    CK(cudaSetDevice(device_));
    lean_ = false;   // everything resident is replaced
    this->N = N;
    A_.N = N;
    if (comm_on_) {   // project_only / ingest run on the host-known count; the slab count is installed afterwards
        A_.comm.on = 0;
    }
    ingest(N, uv, heading, vid, r3d, ids, A_.cur);
    if (project) {
        Launch<R>::project_only(A_, stream_);
        launches_++;
    }
    if (comm_on_) {
        A_.comm.on = 1;
        DevCommState st{N, N};
        CK(cudaMemcpyAsync(comm_state_.p, &st, sizeof(st), cudaMemcpyHostToDevice, stream_));
        CK(cudaMemsetAsync(d_count_.p, 0, d_count_.n * sizeof(int), stream_));
        halo_valid_ = false;   // the first step starts with a pack + exchange + sort of the uploaded particles
        sorted_ = false;
    } else {
        resort(false);
    }
    CK(cudaStreamSynchronize(stream_));
    CK(cudaGetLastError());
    return 0;
}

// device-side seeding (io_kernels.cuh k_seed) + the initial projection and sort: no host array is involved
template <typename R> int Engine<R>::seed_particles(int N, uint64_t seed, int mode, uint32_t first_id)
{
    if (N < 0 || N > capacity_) throw CudaError{"particle count exceeds the context's capacity"};
    if (comm_on_) throw CudaError{"t2d_seed_particles: seed before t2d_comm_init, or upload the slab with t2d_set_state"};
    if (mode != 0 && mode != 1) throw CudaError{"t2d_seed_particles: mode must be 0 (uniform in the chart) or 1 (face centres)"};
    CK(cudaSetDevice(device_));
    lean_ = false;
    this->N = N;
    A_.N = N;
    IoLaunch<R>::seed(N, seed, mode, first_id, chart_.F, A_.mesh.tri, A_.cur, stream_);
    Launch<R>::project_only(A_, stream_);
    launches_ += 2;
    resort(false);
    CK(cudaStreamSynchronize(stream_));
    CK(cudaGetLastError());
    return 0;
}

// slab mode: stable compaction of the owned particles (halo copies are skipped): offsets in d_coff_
template <typename R> int Engine<R>::compact_owned()
{
    DevCommState st;
    CK(cudaMemcpyAsync(&st, comm_state_.p, sizeof(st), cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    const int n = st.n;
    IoLaunch<R>::owned_flags(n, A_.cur, d_cflag_.p, stream_);
    launch_scan(d_cflag_.p, d_coff_.p, d_cblk_.p, n, stream_);
    launches_ += 4;
    int owned = 0;
    CK(cudaMemcpyAsync(&owned, d_coff_.p + n, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    resident_ = n;
    this->N = owned;
    return owned;
}

template <typename R> int Engine<R>::owned_count()
{
    CK(cudaSetDevice(device_));
    if (!comm_on_) return this->N;
    if (!sorted_) throw CudaError{"slab mode: step at least once (or call t2d_step(ctx, 0)) before downloading"};
    return compact_owned();
}

template <typename R> int Engine<R>::download(double* uv, int* heading, int* vid, double* r3d, double* rdot, int* color, int* face)
{
    CK(cudaSetDevice(device_));
    materialize();
    int resident = this->N;
    const int* offsets = nullptr;
    if (comm_on_) {
        owned_count();
        resident = resident_;
        offsets = d_coff_.p;
    }
    const size_t N = (size_t)this->N;
    unsigned char* base = d_stage_out_.p;
    size_t off = 0;
    auto take = [&](size_t bytes) -> void* {
        void* d = base + off;
        off += (bytes + 15) & ~(size_t)15;
        return d;
    };
    HostViewOut o{};
    if (uv) o.uv = (double*)take(16 * N);
    if (heading) o.heading = (int*)take(4 * N);
    if (vid) o.vid = (int*)take(4 * N);
    if (r3d) o.r3d = (double*)take(24 * N);
    if (rdot) o.rdot = (double*)take(16 * N);
    if (color) o.color = (int*)take(4 * N);
    if (face) o.face = (int*)take(4 * N);
    IoLaunch<R>::egest(resident, (int)N, offsets, A_.cur, A_.F, A_.new_heading, o, stream_);
    launches_++;
    if (uv) CK(cudaMemcpyAsync(uv, o.uv, 16 * N, cudaMemcpyDeviceToHost, stream_));
    if (heading) CK(cudaMemcpyAsync(heading, o.heading, 4 * N, cudaMemcpyDeviceToHost, stream_));
    if (vid) CK(cudaMemcpyAsync(vid, o.vid, 4 * N, cudaMemcpyDeviceToHost, stream_));
    if (r3d) CK(cudaMemcpyAsync(r3d, o.r3d, 24 * N, cudaMemcpyDeviceToHost, stream_));
    if (rdot) CK(cudaMemcpyAsync(rdot, o.rdot, 16 * N, cudaMemcpyDeviceToHost, stream_));
    if (color) CK(cudaMemcpyAsync(color, o.color, 4 * N, cudaMemcpyDeviceToHost, stream_));
    if (face) CK(cudaMemcpyAsync(face, o.face, 4 * N, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    return 0;
}

template <typename R> int Engine<R>::download_ids(uint32_t* ids)
{
    CK(cudaSetDevice(device_));
    materialize();
    int resident = this->N;
    const int* offsets = nullptr;
    if (comm_on_) {
        owned_count();
        resident = resident_;
        offsets = d_coff_.p;
    }
    const size_t N = (size_t)this->N;
    HostViewOut o{};
    o.ids = (uint32_t*)d_stage_out_.p;
    IoLaunch<R>::egest(resident, (int)N, offsets, A_.cur, A_.F, A_.new_heading, o, stream_);
    launches_++;
    CK(cudaMemcpyAsync(ids, o.ids, 4 * N, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    return 0;
}

// ---- slab mode -----------------------------------------------------------------------------------------
template <typename R>
int Engine<R>::comm_init(int rank, int world, const uint8_t* id, const double* cuts, EngineBase* const* group)
{
    CK(cudaSetDevice(device_));
    materialize();
    if (P_.neigh_mode != T2D_NEIGH_EUCLID) throw CudaError{"slab mode supports the Euclidean criterion only"};
    if (world < 1 || rank < 0 || rank >= world) throw CudaError{"bad rank / world"};
    if (world > T2D_MAX_WORLD) throw CudaError{"at most 16 slabs"};
    if (comm_on_) comm_destroy();
    for (int r = 0; r + 2 < world; ++r)
        if (!(cuts[r] < cuts[r + 1])) throw CudaError{"slab cuts must be ascending"};
    cuts_.assign(cuts, cuts + (world > 1 ? world - 1 : 0));
    const double big = 1e30;
    auto cut = [&](int k) { return k < 0 ? -big : (k >= world - 1 ? big : cuts_[k]); };   // boundary between slab k and k+1
    const double rmax = std::max(2 * P_.sigma, P_.color_factor * P_.sigma);
    A_.comm.rank = rank;
    A_.comm.world = world;
    A_.comm.lo = (R)cut(rank - 1);
    A_.comm.hi = (R)cut(rank);
    A_.comm.lo2 = (R)cut(rank - 2);
    A_.comm.hi2 = (R)cut(rank + 1);
    A_.comm.halo = (R)(rmax * (1.0 + (sizeof(R) == 8 ? 1e-9 : 1e-3)));
    if (world > 1 && (double)A_.comm.hi - (double)A_.comm.lo < 4.0 * rmax && rank > 0 && rank < world - 1)
        throw CudaError{"slab narrower than 4 r_max"};
    // fixed message capacity per direction: migrants are ~0.1 % of a slab per step, the halo strip ~1 % at 2 M
    // particles per GPU (both measured); T2D_FAULT_COMM_OVERFLOW reports a message that did not fit
    A_.comm.mig_cap = std::max(2048, capacity_ / 256);
    A_.comm.ghost_cap = std::max(8192, capacity_ / 32);
    A_.comm.capacity = capacity_;
    msg_bytes_ = Launch<R>::comm_message_bytes(A_.comm.mig_cap, A_.comm.ghost_cap);
    for (int d = 0; d < 2; ++d) {
        comm_send_[d].alloc(msg_bytes_);
        comm_recv_[d].alloc(msg_bytes_);
        // the messages travel at their full fixed size: define the never-filled tails once (initcheck-clean transports)
        CK(cudaMemsetAsync(comm_send_[d].p, 0, msg_bytes_, stream_));
        CK(cudaMemsetAsync(comm_recv_[d].p, 0, msg_bytes_, stream_));
        CK(cudaMemsetAsync(comm_send_[d].p, 0, 16, stream_));
        CK(cudaMemsetAsync(comm_recv_[d].p, 0, 16, stream_));
        A_.comm.send[d] = comm_send_[d].p;
        A_.comm.recv[d] = comm_recv_[d].p;
    }
    for (int k = 0; k < world - 1; ++k) A_.comm.cuts[k] = (R)cuts_[k];
    // index only the cells this rank can hold: its slab + halo strips (+ one more r_max for rounding and cell edges)
    vox_xlo_ = rank > 0 ? cut(rank - 1) - 3.0 * rmax : -1e300;
    vox_xhi_ = rank < world - 1 ? cut(rank) + 3.0 * rmax : 1e300;
    build_vox();
    far_bytes_ = Launch<R>::comm_far_bytes();
    far_send_.alloc(far_bytes_);
    far_recv_.alloc(far_bytes_ * (size_t)world);
    CK(cudaMemsetAsync(far_send_.p, 0, far_bytes_, stream_));
    CK(cudaMemsetAsync(far_recv_.p, 0, far_bytes_ * (size_t)world, stream_));
    CK(cudaMemsetAsync(far_send_.p, 0, 16, stream_));
    CK(cudaMemsetAsync(far_recv_.p, 0, far_bytes_ * (size_t)world, stream_));
    A_.comm.far_send = far_send_.p;
    A_.comm.far_recv = far_recv_.p;
    A_.comm.far_bytes = far_bytes_;
    comm_state_.alloc(1);
    DevCommState st{this->N, this->N};
    CK(cudaMemcpyAsync(comm_state_.p, &st, sizeof(st), cudaMemcpyHostToDevice, stream_));
    A_.comm.state = comm_state_.p;
    d_cflag_.alloc((size_t)capacity_ + 8);
    d_coff_.alloc((size_t)capacity_ + 8);
    d_cblk_.alloc((size_t)scan_blocks(capacity_) + 8);
    CK(cudaEventCreateWithFlags(&ev_sent_, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ev_consumed_, cudaEventDisableTiming));
    CK(cudaStreamSynchronize(stream_));
    group_.clear();
    peer_[0] = peer_[1] = nullptr;
    if (group) {
        group_.assign(group, group + world);
        peer_[0] = rank > 0 ? group[rank - 1] : nullptr;
        peer_[1] = rank + 1 < world ? group[rank + 1] : nullptr;
    }
    if (id) {
        std::string err;
        link_ = nccl_link_create(rank, world, id, &err);
        if (!link_) throw CudaError{err};
    }
    A_.comm.on = 1;
    comm_on_ = true;
    halo_valid_ = false;
    sorted_ = false;
    return 0;
}

template <typename R> int Engine<R>::comm_destroy()
{
    if (!comm_on_) return 0;
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    if (link_) nccl_link_destroy(link_);
    link_ = nullptr;
    if (ev_sent_) cudaEventDestroy(ev_sent_);
    if (ev_consumed_) cudaEventDestroy(ev_consumed_);
    ev_sent_ = ev_consumed_ = nullptr;
    A_.comm.on = 0;
    comm_on_ = false;
    // `cur` still holds owned particles and halo copies interleaved: nothing resident survives the switch back to
    // single-context mode — the caller uploads a fresh state (t2d_set_state / t2d_set_particles) before stepping
    this->N = 0;
    A_.N = 0;
    halo_valid_ = false;
    if (!closing_ && (vox_xlo_ > -1e299 || vox_xhi_ < 1e299)) {   // back to the index of the whole surface
        vox_xlo_ = -1e300;
        vox_xhi_ = 1e300;
        build_vox();
        sorted_ = false;
    }
    return 0;
}

// advance one step (unless the halo has not been built yet: first call after an upload) and pack the messages
template <typename R> void Engine<R>::comm_phase1()
{
    CK(cudaSetDevice(device_));
    for (int d = 0; d < 2; ++d) CK(cudaMemsetAsync(comm_send_[d].p, 0, 16, stream_));
    CK(cudaMemsetAsync(far_send_.p, 0, 16, stream_));
    prof_mark();
    A_.lean = 0;   // slab mode: the step kernel keeps pos / uv / key (the classification and the messages are built from them)
    if (halo_valid_) {
        A_.step = (uint64_t)step_index;
        if (use_fast2_ && Launch<R>::step_fast2(A_, true, sm_count_, stream_))
            A_.queue_flip++;
        else
            Launch<R>::step_euclid(A_, true, stream_);   // cur -> alt; classifies + packs every particle it has just moved
        std::swap(A_.cur, A_.alt);
        launches_++;
        prof_mark();
    } else {   // first exchange after an upload: classify + pack the resident state as it is
        prof_mark();
        Launch<R>::comm_pack(A_, stream_);
        launches_++;
    }
    prof_mark();
}

template <typename R> void Engine<R>::comm_local_send()
{
    CK(cudaSetDevice(device_));
    const int me = A_.comm.rank, world = A_.comm.world;
    for (int r = 0; r < world; ++r)
        if (r != me) CK(cudaStreamWaitEvent(stream_, group_[r]->comm_event(1), 0));   // rank r has consumed the previous messages
    for (int d = 0; d < 2; ++d) {
        EngineBase* p = peer_[d];
        if (!p) continue;
        CK(cudaMemcpyPeerAsync(p->comm_recv_buffer(1 - d), p->device(), comm_send_[d].p, device_, msg_bytes_, stream_));
    }
    for (int r = 0; r < world; ++r)
        if (r != me) CK(cudaMemcpyPeerAsync(group_[r]->comm_far_slot(me), group_[r]->device(), far_send_.p, device_, far_bytes_, stream_));
    CK(cudaEventRecord(ev_sent_, stream_));
}

template <typename R> void Engine<R>::comm_phase2()
{
    CK(cudaSetDevice(device_));
    if (link_) {
        std::string err;
        if (nccl_exchange(link_, comm_send_[0].p, comm_recv_[0].p, comm_send_[1].p, comm_recv_[1].p, msg_bytes_, far_send_.p,
                          far_recv_.p, far_bytes_, stream_, &err) != 0)
            throw CudaError{err};
    } else {
        for (int r = 0; r < A_.comm.world; ++r)
            if (r != A_.comm.rank) CK(cudaStreamWaitEvent(stream_, group_[r]->comm_event(0), 0));
    }
    Launch<R>::comm_unpack(A_, stream_);
    Launch<R>::comm_unpack_far(A_, stream_);
    if (A_.comm.world > 1) launches_++;
    prof_mark();
    scan_buckets();
    prof_mark();
    if (use_fast2_ && lean_ok_) {   // lean sort: record + aux + source index; materialize() rebuilds the rest on demand
        Launch<R>::scatter_lean(A_, stream_);
        lean_ = true;
    } else {
        l2_window(A_.alt.pos);
        Launch<R>::scatter(A_, stream_);
        lean_ = false;
    }
    std::swap(A_.cur, A_.alt);
    launches_ += 2;
    prof_mark();
    CK(cudaEventRecord(ev_consumed_, stream_));
    if (halo_valid_) {
        step_index++;
        steps_++;
    }
    halo_valid_ = true;
    sorted_ = true;
}

template <typename R> int Engine<R>::comm_finish()
{
    CK(cudaSetDevice(device_));
    CK(cudaStreamSynchronize(stream_));
    CK(cudaGetLastError());
    return read_fault();
}

template <typename R> void Engine<R>::one_step(bool moving, cudaEvent_t* ev, int* nev)
{
    auto mark = [&]() {
        if (ev) CK(cudaEventRecord(ev[(*nev)++], stream_));
    };
    if (!sorted_) resort(false);
    A_.step = (uint64_t)step_index;
    mark();
    if (P_.neigh_mode == T2D_NEIGH_EUCLID) {
#ifdef T2D_F2_ABLATE
        if (const char* e = getenv("T2D_F2_ABLATE_AFTER")) {   // dev builds only: ablation timing, see step_fast2.cuh / tools/gpu_ablate.sh
            static int n_launch = 0;
            const char* m = getenv("T2D_F2_ABLATE_MODE");
            A_.ablate = (++n_launch > atoi(e)) ? (m ? atoi(m) : 0) : 0;
        }
#endif
        // lean pipeline (fp32 fast path, single context): the step kernel writes records, the sort moves record + aux only;
        // pos / uv / r_dot / colour are rebuilt by materialize() when somebody asks for them
        const bool lean = moving && use_fast2_ && lean_ok_ && !comm_on_;
        if (!moving) materialize();   // t2d_forces reports through the full state
        A_.lean = lean ? 1 : 0;
        bool ran_fast2 = false;
        if (use_fast2_ && Launch<R>::step_fast2(A_, moving, sm_count_, stream_)) {
            A_.queue_flip++;
            ran_fast2 = true;
        } else {
            materialize();
            A_.lean = 0;
            Launch<R>::step_euclid(A_, moving, stream_);   // cur -> alt (+ next keys)
        }
        launches_++;
        mark();
        if (moving) std::swap(A_.cur, A_.alt);
        if (moving && lean && ran_fast2) {
            scan_buckets();
            mark();
            Launch<R>::scatter_lean(A_, stream_);
            std::swap(A_.cur, A_.alt);
            launches_++;
            mark();
            lean_ = true;
            step_index++;
            steps_++;
            return;
        }
    } else {
        Launch<R>::neigh_table(A_, stream_, sm_count_);
        launches_++;
        mark();
        if (moving) {
            Launch<R>::wrap_project(A_, stream_);   // in place (+ next keys)
            launches_++;
            mark();
        }
    }
    if (moving) {
        scan_buckets();
        mark();
        l2_window(A_.alt.pos);
        Launch<R>::scatter(A_, stream_);
        std::swap(A_.cur, A_.alt);
        launches_++;
        mark();
        lean_ = false;
        step_index++;
        steps_++;
    }
}

template <typename R> int Engine<R>::read_fault()
{
    DevCounters h;
    CK(cudaMemcpyAsync(&h, d_counters_.p, sizeof(h), cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    int fault = (int)h.fault;
    if (fault) {
        unsigned zero = 0;
        CK(cudaMemcpyAsync(&d_counters_.p->fault, &zero, sizeof(zero), cudaMemcpyHostToDevice, stream_));
        CK(cudaStreamSynchronize(stream_));
    }
    return fault;
}

template <typename R> int Engine<R>::step(int nsteps)
{
    CK(cudaSetDevice(device_));
    if (comm_on_) {
        if (!link_) throw CudaError{"local slab groups are stepped with t2d_step_local"};
        CK(cudaEventRecord(ev0_, stream_));
        if (!halo_valid_) {   // build the halo of the uploaded particles first (no time step)
            comm_phase1();
            comm_phase2();
        }
        for (int s = 0; s < nsteps; ++s) {
            comm_phase1();
            comm_phase2();
        }
        CK(cudaEventRecord(ev1_, stream_));
        CK(cudaStreamSynchronize(stream_));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ev0_, ev1_));
        last_step_ms = ms;
        CK(cudaGetLastError());
        return read_fault();
    }
    if (this->N == 0 || nsteps <= 0) return 0;
    CK(cudaEventRecord(ev0_, stream_));
    for (int s = 0; s < nsteps; ++s) one_step(true, nullptr, nullptr);
    CK(cudaEventRecord(ev1_, stream_));
    CK(cudaStreamSynchronize(stream_));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ev0_, ev1_));
    last_step_ms = ms;
    CK(cudaGetLastError());
    return read_fault();
}

template <typename R>
int Engine<R>::step_host(int N, double* uv, int* heading, int* vid, double* r3d, double* rdot, int* color, bool reproject)
{
    if (comm_on_) throw CudaError{"t2d_step_host is not available in slab mode"};
    {
        // fp32 contexts: float staging + conversion on host threads (step_host32) — when this process has host cores for it.
        // One process per GPU under torchrun: LOCAL_WORLD_SIZE ranks share the host's cores; below 4 cores per rank the plain
        // path (doubles over PCIe, conversion on the device) is faster.  T2D_HOST32 = 0 / 1 forces either path.
        const char* e = getenv("T2D_HOST32");
        const char* lws = getenv("LOCAL_WORLD_SIZE");
        const unsigned per_rank = std::thread::hardware_concurrency() / (unsigned)std::max(1, lws ? atoi(lws) : 1);
        const bool want = e ? atoi(e) != 0 : per_rank >= 4;
        if (sizeof(R) == 4 && N >= 4096 && want) return step_host32(N, uv, heading, vid, r3d, rdot, color, reproject);
    }
    if (reproject)
        set_state(N, uv, heading, nullptr, nullptr, nullptr, true);
    else
        set_state(N, uv, heading, vid, r3d, nullptr, false);
    int fault = step(1);
    download(uv, heading, vid, r3d, rdot, color, nullptr);
    return fault;
}

// ---- asynchronous export (include/t2d.h t2d_export_begin / t2d_export_wait) ------------------------------------------
// layout of one slot (device staging and pinned host buffer alike): uv [2N] double, r3d [3N], rdot [2N], then heading, vid,
// colour [N] int each
template <typename R> int Engine<R>::export_begin(int* slot)
{
    if (comm_on_) throw CudaError{"t2d_export_begin is not available in slab mode"};
    CK(cudaSetDevice(device_));
    const size_t C = (size_t)capacity_, bytes = C * (7 * sizeof(double) + 3 * sizeof(int)) + 64;
    if (!export_stream_) {
        CK(cudaStreamCreateWithFlags(&export_stream_, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) {
            CK(cudaHostAlloc((void**)&h_export_[k], bytes, cudaHostAllocDefault));
            d_export_[k].alloc(bytes);
            CK(cudaEventCreateWithFlags(&ev_export_snap_[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ev_export_done_[k], cudaEventDisableTiming));
            CK(cudaEventRecord(ev_export_done_[k], export_stream_));
        }
    }
    const int k = (int)(export_seq_++ & 1u);
    const size_t N = (size_t)this->N;
    materialize();
    CK(cudaStreamWaitEvent(stream_, ev_export_done_[k], 0));   // the slot's previous copy has left the staging buffer
    unsigned char* base = d_export_[k].p;
    HostViewOut o{};
    o.uv = (double*)base;
    o.r3d = o.uv + 2 * N;
    o.rdot = o.r3d + 3 * N;
    o.heading = (int*)(o.rdot + 2 * N);
    o.vid = o.heading + N;
    o.color = o.vid + N;
    IoLaunch<R>::egest((int)N, (int)N, nullptr, A_.cur, A_.F, A_.new_heading, o, stream_);
    launches_++;
    CK(cudaEventRecord(ev_export_snap_[k], stream_));
    CK(cudaStreamWaitEvent(export_stream_, ev_export_snap_[k], 0));
    CK(cudaMemcpyAsync(h_export_[k], base, N * (7 * sizeof(double) + 3 * sizeof(int)), cudaMemcpyDeviceToHost, export_stream_));
    CK(cudaEventRecord(ev_export_done_[k], export_stream_));
    export_n_[k] = (int)N;
    export_step_[k] = step_index;
    if (slot) *slot = k;
    return 0;
}

template <typename R>
int Engine<R>::export_wait(int slot, int* N, long long* step, const double** uv, const int** heading, const int** vid,
                           const double** r3d, const double** rdot, const int** color)
{
    if (slot < 0 || slot > 1 || !export_stream_) throw CudaError{"t2d_export_wait: no such export in flight"};
    CK(cudaSetDevice(device_));
    CK(cudaEventSynchronize(ev_export_done_[slot]));
    const size_t n = (size_t)export_n_[slot];
    const double* base = (const double*)h_export_[slot];
    if (N) *N = (int)n;
    if (step) *step = export_step_[slot];
    if (uv) *uv = base;
    if (r3d) *r3d = base + 2 * n;
    if (rdot) *rdot = base + 5 * n;
    const int* ib = (const int*)(base + 7 * n);
    if (heading) *heading = ib;
    if (vid) *vid = ib + n;
    if (color) *color = ib + 2 * n;
    return 0;
}

// run fn(t) for t = 0 .. nt-1 on host threads (OpenMP keeps its pool alive between calls; num_threads overrides the
// OMP_NUM_THREADS=1 that torchrun exports)
static void host_parallel(int nt, const std::function<void(int)>& fn)
{
#pragma omp parallel num_threads(nt)
    {
#pragma omp for schedule(static, 1)
        for (int t = 0; t < nt; ++t) fn(t);
    }
}

// t2d_step_host / t2d_step_host_uv for fp32 contexts.  The fast path computes in float, so doubles on the PCIe bus are
// padding: 40 + 136 MB per step at 2 M particles as doubles, 24 + 80 MB as floats.  The caller's arrays stay doubles in the
// reference's layouts; they are narrowed (inputs) and widened (outputs — exact) by host threads in chunks, each chunk while
// the next one is on the bus.  One stream, no synchronisation until the first output chunk is needed.
template <typename R>
int Engine<R>::step_host32(int N, double* uv, int* heading, int* vid, double* r3d, double* rdot, int* color, bool reproject)
{
    if (N < 0 || N > capacity_) throw CudaError{"particle count exceeds the context's capacity"};
    if (!uv || !heading || !vid || !r3d || !rdot || !color) throw CudaError{"t2d_step_host: all six arrays are required"};
    CK(cudaSetDevice(device_));
    const size_t n = (size_t)N;
    if (h32_cap_ < n) {
        if (h32_) cudaFreeHost(h32_);
    for (int k = 0; k < 2; ++k) {
        if (h_export_[k]) cudaFreeHost(h_export_[k]);
        if (ev_export_snap_[k]) cudaEventDestroy(ev_export_snap_[k]);
        if (ev_export_done_[k]) cudaEventDestroy(ev_export_done_[k]);
    }
    if (export_stream_) cudaStreamDestroy(export_stream_);
        h32_ = nullptr;
        CK(cudaHostAlloc((void**)&h32_, sizeof(float) * 12 * (size_t)capacity_ + sizeof(DevCounters) + 64, cudaHostAllocDefault));
        h32_cap_ = (size_t)capacity_;
        for (auto& e : ev_chunk_)
            if (!e) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const char* lws = getenv("LOCAL_WORLD_SIZE");
    int nt = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency() / (unsigned)std::max(1, lws ? atoi(lws) : 1)));
    int K = 4;                                            // chunks per array (<= 16)
    if (const char* e = getenv("T2D_HOST32_T")) nt = std::max(1, std::min(64, atoi(e)));
    if (const char* e = getenv("T2D_HOST32_K")) K = std::max(1, std::min(16, atoi(e)));
    const size_t cs = ((n + K - 1) / K + 63) & ~(size_t)63;
    float* hin = h32_;                                    // [2n] uv, then [3n] r3d
    float* hout = h32_ + 5 * h32_cap_;                    // [2n] uv, [3n] r3d, [2n] rdot
    float* din = (float*)d_stage_in_.p;                   // device staging, same layout as hin + ints behind
    int* din_i = (int*)(din + 5 * n);                     // heading [n], vid [n]
    float* dout = (float*)d_stage_out_.p;                 // [7n] floats, then heading, vid, color
    int* dout_i = (int*)(dout + 7 * n);
    const int ncol_in = reproject ? 2 : 5;
    const bool trace = getenv("T2D_HOST32_TRACE") != nullptr;   // dev: host-side timeline of the phases, microseconds
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count(); };
    double t_in = 0, t_enq = 0, t_first = 0;
    // ---- inputs: narrow a chunk on the host threads, put it on the bus, narrow the next one meanwhile ----
    for (int c = 0; c < K; ++c) {
        const size_t a = std::min(n, (size_t)c * cs), b = std::min(n, a + cs);
        if (a == b) break;
        host_parallel(nt, [&](int t) {
            const size_t len = b - a, ta = a + len * t / nt, tb = a + len * (t + 1) / nt;
            for (int col = 0; col < ncol_in; ++col) {
                const double* src = col < 2 ? uv + col * n : r3d + (col - 2) * n;
                float* dst = hin + col * n;
                for (size_t i = ta; i < tb; ++i) dst[i] = (float)src[i];
            }
        });
        // the chunk of every column in ONE call: `ncol_in` rows of (b - a) floats, row pitch = one column
        CK(cudaMemcpy2DAsync(din + a, sizeof(float) * n, hin + a, sizeof(float) * n, sizeof(float) * (b - a), (size_t)ncol_in,
                             cudaMemcpyHostToDevice, stream_));
    }
    CK(cudaMemcpyAsync(din_i, heading, sizeof(int) * n, cudaMemcpyHostToDevice, stream_));
    if (!reproject) CK(cudaMemcpyAsync(din_i + n, vid, sizeof(int) * n, cudaMemcpyHostToDevice, stream_));
    t_in = since();
    lean_ = false;
    this->N = N;
    A_.N = N;
    if (d_face_hint_.n < (size_t)capacity_) {
        d_face_hint_.alloc((size_t)capacity_);
        face_hint_n_ = -1;
    }
    // re-projection: the caller normally hands back the uv of the previous call, so every particle is still in the same face
    HostViewIn32 in{din, din_i, reproject ? nullptr : din_i + n, reproject ? nullptr : din + 2 * n,
                    (reproject && face_hint_n_ == N) ? d_face_hint_.p : nullptr};
    IoLaunch<R>::ingest32(N, in, A_.cur, stream_);
    launches_++;
    if (reproject) {
        Launch<R>::project_only(A_, stream_);
        launches_++;
    }
    HostViewOut32 o{dout, dout_i, dout_i + n, dout + 2 * n, dout + 5 * n, dout_i + 2 * n, d_face_hint_.p};
    face_hint_n_ = N;
    if (use_fast2_ && lean_ok_ && N > 0) {
        // lean all the way: the upload is sorted straight into records (one full sector per particle instead of six partial
        // ones — the particles arrive in caller order, so this sort is a random permutation), the step runs on them, and
        // the export gathers by caller index through the inverse map the last sort left behind
        if (d_inv_.n < (size_t)capacity_) d_inv_.alloc((size_t)capacity_);
        Launch<R>::bin(A_, stream_);
        scan_buckets();
        A_.lean = 0;
        A_.inv = nullptr;
        Launch<R>::scatter_lean(A_, stream_);
        std::swap(A_.cur, A_.alt);
        launches_ += 2;
        sorted_ = true;
        lean_ = true;
        A_.inv = d_inv_.p;
        one_step(true, nullptr, nullptr);
        A_.inv = nullptr;
        IoLaunch<R>::egest32_lean(N, d_inv_.p, A_.src, A_.cur, A_.alt, o, stream_);
        launches_++;
    } else {
        resort(false);
        if (N > 0) one_step(true, nullptr, nullptr);
        materialize();
        IoLaunch<R>::egest32(N, A_.cur, o, stream_);
        launches_++;
    }
    // ---- outputs: chunk c of every array goes on the bus, an event marks it; the host widens chunk c while c+1 travels ----
    // the counters land in PINNED memory: a copy into pageable memory would block the host here until the whole step is done
    DevCounters& hc = *reinterpret_cast<DevCounters*>(reinterpret_cast<unsigned char*>(h32_) + sizeof(float) * 12 * h32_cap_ + 16);
    hc.fault = 0;
    int nchunks = 0;
    for (int c = 0; c < K; ++c) {
        const size_t a = std::min(n, (size_t)c * cs), b = std::min(n, a + cs);
        if (a == b) break;
        CK(cudaMemcpy2DAsync(hout + a, sizeof(float) * n, dout + a, sizeof(float) * n, sizeof(float) * (b - a), 7,
                             cudaMemcpyDeviceToHost, stream_));   // 7 float columns of the chunk in one call
        CK(cudaMemcpyAsync(heading + a, dout_i + a, sizeof(int) * (b - a), cudaMemcpyDeviceToHost, stream_));
        CK(cudaMemcpyAsync(vid + a, dout_i + n + a, sizeof(int) * (b - a), cudaMemcpyDeviceToHost, stream_));
        CK(cudaMemcpyAsync(color + a, dout_i + 2 * n + a, sizeof(int) * (b - a), cudaMemcpyDeviceToHost, stream_));
        if (c == 0) CK(cudaMemcpyAsync(&hc, d_counters_.p, sizeof(hc), cudaMemcpyDeviceToHost, stream_));
        CK(cudaEventRecord(ev_chunk_[c], stream_));
        nchunks++;
    }
    t_enq = since();
    for (int c = 0; c < nchunks; ++c) {
        const size_t a = std::min(n, (size_t)c * cs), b = std::min(n, a + cs);
        CK(cudaEventSynchronize(ev_chunk_[c]));
        if (c == 0) t_first = since();
        host_parallel(nt, [&](int t) {
            const size_t len = b - a, ta = a + len * t / nt, tb = a + len * (t + 1) / nt;
            for (int col = 0; col < 7; ++col) {
                double* dst = col < 2 ? uv + col * n : (col < 5 ? r3d + (col - 2) * n : rdot + (col - 5) * n);
                const float* src = hout + col * n;
                size_t i = ta;
#if defined(__SSE2__)
                // streaming stores: the caller's arrays are written once and not read here — no write-allocate traffic
                for (; i < tb && (((uintptr_t)(dst + i)) & 15); ++i) dst[i] = (double)src[i];
                for (; i + 2 <= tb; i += 2) _mm_stream_pd(dst + i, _mm_set_pd((double)src[i + 1], (double)src[i]));
#endif
                for (; i < tb; ++i) dst[i] = (double)src[i];
            }
        });
    }
    CK(cudaStreamSynchronize(stream_));
    CK(cudaGetLastError());
    if (trace)
        fprintf(stderr, "t2d host32: inputs narrowed + enqueued %.0f us, everything enqueued %.0f us, first output chunk landed %.0f us, "
                        "done %.0f us\n", t_in, t_enq, t_first, since());
    int fault = N > 0 ? (int)hc.fault : 0;
    if (fault) {
        unsigned zero = 0;
        CK(cudaMemcpyAsync(&d_counters_.p->fault, &zero, sizeof(zero), cudaMemcpyHostToDevice, stream_));
        CK(cudaStreamSynchronize(stream_));
    }
    return fault;
}

template <typename R> int Engine<R>::forces(double* F, int* new_heading, int* color)
{
    if (comm_on_) throw CudaError{"t2d_forces is not available in slab mode"};
    CK(cudaSetDevice(device_));
    if (this->N == 0) return 0;
    A_.write_F = 1;
    one_step(false, nullptr, nullptr);
    A_.write_F = 0;
    const size_t N = (size_t)this->N;
    unsigned char* base = d_stage_out_.p;
    HostViewOut o{};
    size_t off = 0;
    o.F = (double*)(base + off);
    off += 16 * N;
    o.new_heading = (int*)(base + off);
    off += (4 * N + 15) & ~(size_t)15;
    o.color = (int*)(base + off);
    ParticleArrays<R> view = A_.cur;
    if (P_.neigh_mode == T2D_NEIGH_EUCLID) view.color = A_.alt.color;   // the Euclid kernels report into the scratch side
    IoLaunch<R>::egest((int)N, (int)N, nullptr, view, A_.F, A_.new_heading, o, stream_);
    launches_++;
    if (F) CK(cudaMemcpyAsync(F, o.F, 16 * N, cudaMemcpyDeviceToHost, stream_));
    if (new_heading) CK(cudaMemcpyAsync(new_heading, o.new_heading, 4 * N, cudaMemcpyDeviceToHost, stream_));
    if (color) CK(cudaMemcpyAsync(color, o.color, 4 * N, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    CK(cudaGetLastError());
    return 0;
}

template <typename R> int Engine<R>::observables(double* out)
{
    CK(cudaSetDevice(device_));
    materialize();
    launch_observables(A_.cur.pos, A_.cur.rdot, comm_on_ ? A_.cur.aux : nullptr, sizeof(R) == 4, comm_on_ ? capacity_ : this->N,
                       comm_on_ ? &comm_state_.p->n : nullptr, d_trig_d_.p, d_obs_.p, stream_);
    launches_++;
    double h[T2D_OBS_LEN];
    CK(cudaMemcpyAsync(h, d_obs_.p, sizeof(h), cudaMemcpyDeviceToHost, stream_));
    DevCounters c;
    CK(cudaMemcpyAsync(&c, d_counters_.p, sizeof(c), cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    const double n = h[T2D_OBS_COUNT];   // owned particles (slab mode skips the halo copies)
    h[T2D_OBS_PHI] = n > 0 ? sqrt(h[T2D_OBS_SUM_COS] * h[T2D_OBS_SUM_COS] + h[T2D_OBS_SUM_SIN] * h[T2D_OBS_SUM_SIN]) / n : 0;
    h[T2D_OBS_MEAN_SPEED] = n > 0 ? h[T2D_OBS_SUM_SPEED] / n : 0;
    h[T2D_OBS_LOST] = (double)c.lost;
    h[T2D_OBS_NONFINITE] = (double)c.nonfinite;
    memcpy(out, h, sizeof(h));
    return 0;
}

template <typename R> int Engine<R>::get_counters(t2d_counters* out)
{
    CK(cudaSetDevice(device_));
    DevCounters c;
    CK(cudaMemcpyAsync(&c, d_counters_.p, sizeof(c), cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    memset(out, 0, sizeof(*out));
    out->steps = steps_;
    out->kernel_launches = launches_;
    out->pairs_in_range = (int64_t)c.pairs_in_range;
    out->ties_cutoff = (int64_t)c.ties_cutoff;
    out->ties_trunc = (int64_t)c.ties_trunc;
    out->wraps = (int64_t)c.wraps;
    out->wrap_cap_hits = (int64_t)c.wrap_cap_hits;
    out->order_fallbacks = (int64_t)c.order_fallbacks;
    out->trig_fallbacks = (int64_t)c.trig_fallbacks;
    out->locate_fallbacks = (int64_t)c.locate_fallbacks;
    out->max_row = (int64_t)c.max_row;
    out->cell_fallbacks = (int64_t)c.cell_fallbacks;
    out->buckets = A_.M;
    return 0;
}

template <typename R> int Engine<R>::reset_counters()
{
    CK(cudaSetDevice(device_));
    CK(cudaMemsetAsync(d_counters_.p, 0, sizeof(DevCounters), stream_));
    CK(cudaStreamSynchronize(stream_));
    steps_ = 0;
    launches_ = 0;
    return 0;
}

// CellHelper::get_r3d on caller points, using the spare half of the double buffer as scratch
template <typename R> int Engine<R>::get_r3d(int N, const double* uv, double* r3d, int* vid, int* face)
{
    if (N < 0 || N > capacity_) throw CudaError{"N exceeds the context's capacity"};
    CK(cudaSetDevice(device_));
    materialize();
    StepArgs<R> T = A_;
    T.N = N;
    T.comm.on = 0;
    T.cur = A_.alt;
    ingest(N, uv, nullptr, nullptr, nullptr, nullptr, T.cur);
    Launch<R>::project_only(T, stream_);
    launches_++;
    HostViewOut o{};
    unsigned char* base = d_stage_out_.p;
    o.r3d = (double*)base;
    o.vid = (int*)(base + 24 * (size_t)N);
    o.face = (int*)(base + 24 * (size_t)N + ((4 * (size_t)N + 15) & ~(size_t)15));
    IoLaunch<R>::egest(N, N, nullptr, T.cur, A_.F, A_.new_heading, o, stream_);
    launches_++;
    if (r3d) CK(cudaMemcpyAsync(r3d, o.r3d, 24 * (size_t)N, cudaMemcpyDeviceToHost, stream_));
    if (vid) CK(cudaMemcpyAsync(vid, o.vid, 4 * (size_t)N, cudaMemcpyDeviceToHost, stream_));
    if (face) CK(cudaMemcpyAsync(face, o.face, 4 * (size_t)N, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    CK(cudaGetLastError());
    return 0;
}

template <typename R> int Engine<R>::tiling(int N, double* uv_old, double* uv, int* heading)
{
    if (N < 0 || N > capacity_) throw CudaError{"N exceeds the context's capacity"};
    CK(cudaSetDevice(device_));
    materialize();
    unsigned char* base = d_stage_in_.p;
    double* s_old = (double*)base;
    double* s_new = (double*)(base + 16 * (size_t)N);
    CK(cudaMemcpyAsync(s_old, uv_old, 16 * (size_t)N, cudaMemcpyHostToDevice, stream_));
    CK(cudaMemcpyAsync(s_new, uv, 16 * (size_t)N, cudaMemcpyHostToDevice, stream_));
    CK(cudaMemcpyAsync(A_.new_heading, heading, 4 * (size_t)N, cudaMemcpyHostToDevice, stream_));
    Real2<R>* d_old = A_.alt.uv;
    Real2<R>* d_new = A_.uv_new;
    IoLaunch<R>::in2(N, s_old, d_old, stream_);
    IoLaunch<R>::in2(N, s_new, d_new, stream_);
    Launch<R>::tiling_only(A_, d_old, d_new, A_.new_heading, N, stream_);
    double* o_old = (double*)d_stage_out_.p;
    double* o_new = (double*)(d_stage_out_.p + 16 * (size_t)N);
    IoLaunch<R>::out2(N, d_old, o_old, stream_);
    IoLaunch<R>::out2(N, d_new, o_new, stream_);
    launches_ += 5;
    CK(cudaMemcpyAsync(uv_old, o_old, 16 * (size_t)N, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(uv, o_new, 16 * (size_t)N, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(heading, A_.new_heading, 4 * (size_t)N, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    CK(cudaGetLastError());
    return read_fault();
}

template <typename R> int Engine<R>::unit_vectors(int N, const int* heading, double* out)
{
    if (N < 0 || N > capacity_) throw CudaError{"N exceeds the context's capacity"};
    CK(cudaSetDevice(device_));
    CK(cudaMemcpyAsync(A_.new_heading, heading, 4 * (size_t)N, cudaMemcpyHostToDevice, stream_));
    R* tmp = reinterpret_cast<R*>(A_.uv_new);
    Launch<R>::unit_vectors(A_, A_.new_heading, tmp, N, stream_);
    double* o = (double*)d_stage_out_.p;
    IoLaunch<R>::outN(N, 2, tmp, o, stream_);
    launches_ += 2;
    CK(cudaMemcpyAsync(out, o, 16 * (size_t)N, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    CK(cudaGetLastError());
    return 0;
}

template <typename R> int Engine<R>::hop_table(uint8_t* out)
{
    CK(cudaSetDevice(device_));
    std::vector<uint8_t> h;
    build_hop_table_device(&h);
    memcpy(out, h.data(), h.size());
    return 0;
}

template <typename R> int Engine<R>::profile_step(const char** names, double* ms, int cap)
{
    CK(cudaSetDevice(device_));
    if (comm_on_) {   // collective in NCCL mode: every rank must call it
        if (!link_) throw CudaError{"t2d_profile_step: local slab groups are not supported"};
        static const char* kSlab[] = {"step_fused", "comm_pack", "exchange_unpack", "scan", "scatter"};
        if (!halo_valid_) step(0);
        cudaEvent_t ev[8];
        for (auto& e : ev) CK(cudaEventCreate(&e));
        int nev = 0;
        prof_ev_ = ev;
        prof_nev_ = &nev;
        comm_phase1();
        comm_phase2();
        prof_ev_ = nullptr;
        prof_nev_ = nullptr;
        CK(cudaStreamSynchronize(stream_));
        int n = std::min(cap, nev - 1);
        for (int i = 0; i < n; ++i) {
            float t = 0;
            CK(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
            names[i] = kSlab[i];
            ms[i] = t;
        }
        for (auto& e : ev) cudaEventDestroy(e);
        read_fault();
        return n;
    }
    if (this->N == 0) return 0;
    static const char* kEuclid[] = {"step_fused", "scan", "scatter"};
    static const char* kTable[] = {"neigh_table", "wrap_project", "scan", "scatter"};
    const char** kNames = P_.neigh_mode == T2D_NEIGH_EUCLID ? kEuclid : kTable;
    if (!sorted_) resort(false);
    cudaEvent_t ev[8];
    for (auto& e : ev) CK(cudaEventCreate(&e));
    int nev = 0;
    one_step(true, ev, &nev);
    CK(cudaStreamSynchronize(stream_));
    int n = std::min(cap, nev - 1);
    for (int i = 0; i < n; ++i) {
        float t = 0;
        CK(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
        names[i] = kNames[i];
        ms[i] = t;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    read_fault();
    return n;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------------------------------
struct t2d_ctx {
    std::string err;
    std::unique_ptr<EngineBase> eng;
};

#define T2D_TRY(ctx, body)                          \
    try {                                           \
        body                                        \
    } catch (const CudaError& e) {                  \
        (ctx)->err = e.msg;                         \
        return -1;                                  \
    } catch (const std::exception& e) {             \
        (ctx)->err = e.what();                      \
        return -1;                                  \
    }

extern "C" {

int t2d_version(void) { return 100; }

int t2d_create(const t2d_mesh* mesh, const t2d_table* table, const t2d_params* params, int device, t2d_ctx** out)
{
    if (!out || !params) {
        g_create_error = "null argument";
        return -1;
    }
    *out = nullptr;
    std::unique_ptr<t2d_ctx> c(new t2d_ctx());
    try {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                             " — lib2dtissue_b200 has no CPU fallback";
            return -1;
        }
        if (params->precision == T2D_PRECISION_FP32)
            c->eng.reset(new Engine<float>(mesh, table, params, device));
        else
            c->eng.reset(new Engine<double>(mesh, table, params, device));
    } catch (const CudaError& e) {
        g_create_error = e.msg;
        return -1;
    } catch (const std::exception& e) {
        g_create_error = e.what();
        return -1;
    }
    *out = c.release();
    return 0;
}

void t2d_destroy(t2d_ctx* ctx) { delete ctx; }

const char* t2d_last_error(const t2d_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int t2d_set_particles(t2d_ctx* ctx, int32_t N, const double* uv, const int32_t* heading, const uint32_t* ids)
{
    T2D_TRY(ctx, return ctx->eng->set_state(N, uv, heading, nullptr, nullptr, ids, true);)
}
int t2d_set_state(t2d_ctx* ctx, int32_t N, const double* uv, const int32_t* heading, const int32_t* vid, const double* r3d,
                  const uint32_t* ids)
{
    T2D_TRY(ctx, return ctx->eng->set_state(N, uv, heading, vid, r3d, ids, false);)
}
int t2d_download(t2d_ctx* ctx, double* uv, int32_t* heading, int32_t* vid, double* r3d, double* rdot, int32_t* color,
                 int32_t* face)
{
    T2D_TRY(ctx, return ctx->eng->download(uv, heading, vid, r3d, rdot, color, face);)
}
int32_t t2d_particle_count(const t2d_ctx* ctx) { return ctx->eng->N; }
int t2d_step(t2d_ctx* ctx, int32_t nsteps) { T2D_TRY(ctx, return ctx->eng->step(nsteps);) }
int t2d_step_host(t2d_ctx* ctx, int32_t N, double* uv, int32_t* heading, int32_t* vid, double* r3d, double* rdot,
                  int32_t* color)
{
    T2D_TRY(ctx, return ctx->eng->step_host(N, uv, heading, vid, r3d, rdot, color, false);)
}
int t2d_step_host_uv(t2d_ctx* ctx, int32_t N, double* uv, int32_t* heading, int32_t* vid_out, double* r3d_out, double* rdot,
                     int32_t* color)
{
    T2D_TRY(ctx, return ctx->eng->step_host(N, uv, heading, vid_out, r3d_out, rdot, color, true);)
}
int t2d_observables(t2d_ctx* ctx, double out[T2D_OBS_LEN]) { T2D_TRY(ctx, return ctx->eng->observables(out);) }
int t2d_get_counters(t2d_ctx* ctx, t2d_counters* out) { T2D_TRY(ctx, return ctx->eng->get_counters(out);) }
int t2d_reset_counters(t2d_ctx* ctx) { T2D_TRY(ctx, return ctx->eng->reset_counters();) }
int t2d_set_tie_log(t2d_ctx* ctx, int on) { T2D_TRY(ctx, return ctx->eng->set_tie_log(on);) }
int t2d_export_begin(t2d_ctx* ctx, int32_t* slot)
{
    int k = 0;
    T2D_TRY(ctx, { int rc = ctx->eng->export_begin(&k); if (slot) *slot = k; return rc; })
}
int t2d_export_wait(t2d_ctx* ctx, int32_t slot, int32_t* N, int64_t* step, const double** uv, const int32_t** heading,
                    const int32_t** vid, const double** r3d, const double** rdot, const int32_t** color)
{
    long long st = 0;
    int n = 0;
    T2D_TRY(ctx, {
        int rc = ctx->eng->export_wait(slot, &n, &st, uv, heading, vid, r3d, rdot, color);
        if (N) *N = n;
        if (step) *step = st;
        return rc;
    })
}
int t2d_seed_particles(t2d_ctx* ctx, int32_t N, uint64_t seed, int32_t mode, uint32_t first_id)
{
    T2D_TRY(ctx, return ctx->eng->seed_particles(N, seed, mode, first_id);)
}
int64_t t2d_get_step(const t2d_ctx* ctx) { return ctx->eng->step_index; }
int t2d_set_step(t2d_ctx* ctx, int64_t step)
{
    ctx->eng->step_index = step;
    return 0;
}
int t2d_set_params(t2d_ctx* ctx, const t2d_params* params) { T2D_TRY(ctx, return ctx->eng->set_params(params);) }
int t2d_get_r3d(t2d_ctx* ctx, int32_t N, const double* uv, double* r3d, int32_t* vid, int32_t* face)
{
    T2D_TRY(ctx, return ctx->eng->get_r3d(N, uv, r3d, vid, face);)
}
int t2d_tiling(t2d_ctx* ctx, int32_t N, double* uv_old, double* uv, int32_t* heading)
{
    T2D_TRY(ctx, return ctx->eng->tiling(N, uv_old, uv, heading);)
}
int t2d_angles_to_unit_vectors(t2d_ctx* ctx, int32_t N, const int32_t* heading, double* out)
{
    T2D_TRY(ctx, return ctx->eng->unit_vectors(N, heading, out);)
}
int t2d_forces(t2d_ctx* ctx, double* F, int32_t* new_heading, int32_t* color)
{
    T2D_TRY(ctx, return ctx->eng->forces(F, new_heading, color);)
}
int t2d_build_hop_table(t2d_ctx* ctx, uint8_t* out) { T2D_TRY(ctx, return ctx->eng->hop_table(out);) }
double t2d_last_step_ms(const t2d_ctx* ctx) { return ctx->eng->last_step_ms; }
int t2d_profile_step(t2d_ctx* ctx, const char** names, double* ms, int cap)
{
    T2D_TRY(ctx, return ctx->eng->profile_step(names, ms, cap);)
}

void* t2d_pinned_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void t2d_pinned_free(void* p)
{
    if (p) cudaFreeHost(p);
}

// ---- multi-GPU slabs --------------------------------------------------------------------------------------
int t2d_comm_unique_id(uint8_t id[T2D_UNIQUE_ID_BYTES])
{
    std::string err;
    if (nccl_unique_id(id, &err) != 0) {
        g_create_error = err;
        return -1;
    }
    return 0;
}
int t2d_comm_init(t2d_ctx* ctx, int rank, int world, const uint8_t id[T2D_UNIQUE_ID_BYTES], const double* cuts)
{
    if (!id) {
        ctx->err = "t2d_comm_init needs the NCCL unique id of rank 0 (t2d_comm_unique_id)";
        return -1;
    }
    T2D_TRY(ctx, return ctx->eng->comm_init(rank, world, id, cuts, nullptr);)
}
int t2d_comm_init_local(t2d_ctx** ctxs, int world, const double* cuts)
{
    std::vector<EngineBase*> group;
    for (int r = 0; r < world; ++r) group.push_back(ctxs[r]->eng.get());
    for (int r = 0; r < world; ++r) {
        std::string msg;
        try {
            ctxs[r]->eng->comm_init(r, world, nullptr, cuts, group.data());
            continue;
        } catch (const CudaError& e) {
            msg = e.msg;
        } catch (const std::exception& e) {
            msg = e.what();
        }
        // the group is all-or-nothing: undo the ranks already initialised, report on ctxs[0] (what callers read)
        for (int q = 0; q < r; ++q) {
            try {
                ctxs[q]->eng->comm_destroy();
            } catch (...) {
            }
        }
        ctxs[r]->err = msg;
        ctxs[0]->err = "rank " + std::to_string(r) + ": " + msg;
        return -1;
    }
    return 0;
}
// lockstep drive of a local group: every phase is enqueued for all ranks before the next one, the streams are
// ordered against each other with events only (no host synchronisation until the end)
int t2d_step_local(t2d_ctx** ctxs, int world, int32_t nsteps)
{
    int fault = 0;
    try {
        for (int r = 0; r < world; ++r)
            if (!ctxs[r]->eng->comm_is_local()) {
                ctxs[r]->err = "context is not part of a local slab group";
                return -1;
            }
        // the first pass after an upload only builds the halo (no time step); comm_phase2 tracks that per context
        for (int s = -1; s < nsteps; ++s) {
            if (s < 0 && ctxs[0]->eng->halo_ready()) continue;
            for (int r = 0; r < world; ++r) ctxs[r]->eng->comm_phase1();
            for (int r = 0; r < world; ++r) ctxs[r]->eng->comm_local_send();
            for (int r = 0; r < world; ++r) ctxs[r]->eng->comm_phase2();
        }
        for (int r = 0; r < world; ++r) fault |= ctxs[r]->eng->comm_finish();
    } catch (const CudaError& e) {
        ctxs[0]->err = e.msg;
        return -1;
    } catch (const std::exception& e) {
        ctxs[0]->err = e.what();
        return -1;
    }
    return fault;
}
int t2d_comm_destroy(t2d_ctx* ctx) { T2D_TRY(ctx, return ctx->eng->comm_destroy();) }
int32_t t2d_owned_count(t2d_ctx* ctx) { T2D_TRY(ctx, return ctx->eng->owned_count();) }
int t2d_download_ids(t2d_ctx* ctx, uint32_t* ids) { T2D_TRY(ctx, return ctx->eng->download_ids(ids);) }

}  // extern "C"
