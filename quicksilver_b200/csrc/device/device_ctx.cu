// device_ctx.cu -- qsb_ctx: one GPU's share of the problem (see include/qsb.h, "Device context").
//
// Owns: the flattened problem image in HBM (geometry arrays + one contiguous "hot block" of nuclear
// data covered by a persisting-L2 access-policy window), two SoA particle vaults (processing, census),
// per-peer send slabs, the scalar-flux array and the control block.  All work is issued on one
// non-blocking stream.  There is no host fallback: any CUDA failure is returned as QSB_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "device_types.cuh"

using namespace qsb;

namespace {

struct CudaFailure { std::string what; };

void check(cudaError_t e, const char* what)
{
    if (e != cudaSuccess) throw CudaFailure{ std::string(what) + ": " + cudaGetErrorString(e) };
}
#define QSB_CUDA(call) check((call), #call)

template <typename T>
T* devAlloc(size_t n, std::vector<void*>& owned)
{
    void* p = nullptr;
    QSB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    owned.push_back(p);
    return static_cast<T*>(p);
}

template <typename T>
T* devUpload(const T* host, size_t n, std::vector<void*>& owned)
{
    T* p = devAlloc<T>(n, owned);
    if (n) QSB_CUDA(cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice));
    return p;
}

// ---- small service kernels -------------------------------------------------------------------------

// host AoS records (MC_Base_Particle layout) -> SoA slots [first, first+n) of the processing vault
// boundary-first list (device_types.cuh, kPrioTicket): the vault slots below n whose cell is near another rank, in no
// particular order (one atomic per warp); count -> ctl->prio_count
__global__ void boundary_list_kernel(const int* __restrict__ cell, unsigned long long n, const uint8_t* __restrict__ cell_near,
                                     uint32_t* __restrict__ list, unsigned long long* __restrict__ count)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long rounded = (n + 31ull) & ~31ull;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < rounded; i += stride)
    {
        const bool hit = i < n && cell_near[cell[i]] != 0;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m == 0u) continue;
        unsigned long long base = 0;
        if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(count, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (hit) list[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
    }
}

__global__ void aos_to_soa_kernel(const qsb_base_particle* __restrict__ in, unsigned long long n, VaultView v,
                                  unsigned long long first, const int* __restrict__ domain_cell_offset, uint32_t epoch)
{
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
    {
        const qsb_base_particle b = in[i];
        const unsigned long long s = first + i;
        v.x[s] = b.coordinate[0]; v.y[s] = b.coordinate[1]; v.z[s] = b.coordinate[2];
        v.vx[s] = b.velocity[0]; v.vy[s] = b.velocity[1]; v.vz[s] = b.velocity[2];
        v.energy[s] = b.kinetic_energy; v.weight[s] = b.weight; v.ttc[s] = b.time_to_census; v.age[s] = b.age;
        v.nmfp[s] = b.num_mean_free_paths; v.nseg[s] = b.num_segments;
        v.seed[s] = b.random_number_seed; v.id[s] = b.identifier;
        v.cell[s] = domain_cell_offset[b.domain] + b.cell;
        v.tags[s] = make_int4(b.last_event, b.num_collisions, b.breed, b.species);
        v.dirx[s] = nan; v.diry[s] = nan; v.dirz[s] = nan;
        v.ready[s] = epoch;
    }
}

// arrivals from other ranks (records already in device memory) -> SoA slots, direction cosine kept
__global__ void arrivals_to_soa_kernel(const ExchangeRecord* __restrict__ in, unsigned long long n, VaultView v,
                                       unsigned long long first, const int* __restrict__ domain_cell_offset, uint32_t epoch)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
    {
        const ExchangeRecord r = in[i];
        const qsb_base_particle& b = r.p;
        const unsigned long long s = first + i;
        v.x[s] = b.coordinate[0]; v.y[s] = b.coordinate[1]; v.z[s] = b.coordinate[2];
        v.vx[s] = b.velocity[0]; v.vy[s] = b.velocity[1]; v.vz[s] = b.velocity[2];
        v.energy[s] = b.kinetic_energy; v.weight[s] = b.weight; v.ttc[s] = b.time_to_census; v.age[s] = b.age;
        v.nmfp[s] = b.num_mean_free_paths; v.nseg[s] = b.num_segments;
        v.seed[s] = b.random_number_seed; v.id[s] = b.identifier;
        v.cell[s] = domain_cell_offset[b.domain] + b.cell;
        v.tags[s] = make_int4(b.last_event, b.num_collisions, b.breed, b.species);
        v.dirx[s] = r.dir[0]; v.diry[s] = r.dir[1]; v.dirz[s] = r.dir[2];
        __threadfence();
        v.ready[s] = epoch;
    }
}

// census vault SoA -> AoS records for the host
__global__ void soa_to_aos_kernel(VaultView v, unsigned long long first, unsigned long long n, qsb_base_particle* __restrict__ out,
                                  const int* __restrict__ domain_cell_offset, int n_domains)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
    {
        const unsigned long long s = first + i;
        qsb_base_particle b;
        b.coordinate[0] = v.x[s]; b.coordinate[1] = v.y[s]; b.coordinate[2] = v.z[s];
        b.velocity[0] = v.vx[s]; b.velocity[1] = v.vy[s]; b.velocity[2] = v.vz[s];
        b.kinetic_energy = v.energy[s]; b.weight = v.weight[s]; b.time_to_census = v.ttc[s]; b.age = v.age[s];
        b.num_mean_free_paths = v.nmfp[s]; b.num_segments = v.nseg[s];
        b.random_number_seed = v.seed[s]; b.identifier = v.id[s];
        const int4 t = v.tags[s];
        b.last_event = t.x; b.num_collisions = t.y; b.breed = t.z; b.species = t.w;
        const int flat = v.cell[s];
        int d = 0;
        while (d + 1 < n_domains && flat >= domain_cell_offset[d + 1]) d++;
        b.domain = d; b.cell = flat - domain_cell_offset[d];
        out[i] = b;
    }
}

// scalar-flux grand total: fixed-shape two-stage tree, so the result is reproducible run to run
__global__ void flux_partial_sum_kernel(const double* __restrict__ flux, unsigned long long n, double* __restrict__ partial)
{
    __shared__ double sh[256];
    double acc = 0.0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        acc += flux[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1)
    {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

} // namespace

// Fluence::compute (src/Tallies.cc:100-121): fluence[cell] += sum over groups of this cycle's scalar flux.  One warp per
// cell, lanes stride the groups, fixed-shape tree: reproducible; reads the flux array once at HBM speed instead of sending
// it to the host (482 MB per cycle at 64^3 cells x 230 groups).
__global__ void fluence_accumulate_kernel(const double* __restrict__ flux, int n_cells, int n_groups, double* __restrict__ fluence)
{
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int cell = blockIdx.x * warps_per_block + (threadIdx.x >> 5); cell < n_cells; cell += gridDim.x * warps_per_block)
    {
        const double* row = flux + (size_t)cell * n_groups;
        double sum = 0.0;
        for (int g = lane; g < n_groups; g += 32) sum += row[g];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        if (lane == 0) fluence[cell] += sum;
    }
}

// EnergySpectrum::UpdateSpectrum (src/EnergySpectrum.cc:12-35) over a census in device memory: energy group of every record
// (NuclearData::getEnergyGroup, src/NuclearData.cc:208-227, the reference's bisection on the same doubles), counted in a
// per-block shared-memory histogram, flushed with one atomic per non-empty bin per block.  `energy` + i * stride_doubles
// addresses record i's kinetic energy: stride 1 for the SoA census vault, 17 for the 136-byte records of a streamed census.
__global__ void census_energy_histogram_kernel(const double* __restrict__ energy, unsigned long long n, int stride_doubles,
                                               const double* __restrict__ edges, int n_edges, unsigned long long* __restrict__ hist)
{
    extern __shared__ unsigned int s_hist[];
    for (int b = threadIdx.x; b < n_edges; b += blockDim.x) s_hist[b] = 0u;
    __syncthreads();
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
    {
        const double e = __ldcs(energy + i * (unsigned long long)stride_doubles);
        int group;
        if (e <= __ldg(edges)) group = 0;
        else if (e > __ldg(edges + n_edges - 1)) group = n_edges - 1;
        else
        {
            int low = 0, high = n_edges - 1;
            while (high != low + 1)
            {
                const int mid = (high + low) / 2;
                if (e < __ldg(edges + mid)) high = mid; else low = mid;
            }
            group = low;
        }
        atomicAdd(&s_hist[group], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n_edges; b += blockDim.x)
        if (s_hist[b]) atomicAdd(&hist[b], (unsigned long long)s_hist[b]);
}

struct qsb_ctx
{
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    qsb_options opt{};
    std::vector<void*> owned;
    DevImage im{};
    int n_ranks = 1, my_rank = 0;
    std::vector<int32_t> host_domain_offset;
    const int* d_domain_offset = nullptr;
    VaultView vault[2]{};          // [proc], [census]
    int proc = 0;
    ExchangeRecord* sends = nullptr;
    unsigned long long send_capacity = 0;
    DevControl* d_ctl = nullptr;
    DevControl* h_ctl = nullptr;   // pinned mirror
    double* flux = nullptr;
    double* fluence = nullptr;     // [n_cells], allocated by the first qsb_fluence_accumulate
    double* flux_partial = nullptr;
    double* h_partial = nullptr;   // pinned
    void* staging = nullptr;       // device AoS staging for put/get
    size_t staging_records = 0;
    void* h_staging = nullptr;     // pinned host staging (two halves)
    double dt = 0;
    unsigned long long host_tail = 0;   // slots written from the host side this cycle
    unsigned long long ready_prefix = 0;
    unsigned long long consumed = 0;         // slots fully processed by earlier qsb_track calls of this cycle
    unsigned long long pending_inflight = 0; // histories written from the host side since the last qsb_track
    // host-buffer streaming (qsb_stream_begin / qsb_track / qsb_stream_end, see include/qsb.h)
    cudaStream_t stream_in = nullptr, stream_out = nullptr;
    cudaEvent_t ev_ctl = nullptr, ev_stage[2] = { nullptr, nullptr };
    cudaEvent_t ev_in_done = nullptr;           // QSB_TRACE: the last input chunk of a streamed cycle has landed (timed against ev0)
    qsb_base_particle* d_in_aos = nullptr;      size_t in_aos_cap = 0;
    qsb_base_particle* d_census_aos = nullptr;  // [vault capacity]
    unsigned int* d_chunk_done = nullptr;       // [n_chunks]
    unsigned int* h_chunk_flags = nullptr;      // [n_chunks] pinned + mapped
    unsigned int* d_chunk_flags = nullptr;      // device alias of h_chunk_flags
    unsigned long long* h_marks = nullptr;      // [n_marks] pinned: values copied into ctl->in_ready, one per input chunk
    size_t n_marks = 0;
    size_t n_chunks = 0;
    bool streaming = false, stream_input_issued = false;
    uint32_t* d_prio_list = nullptr; size_t prio_list_cap = 0;   // boundary-first list (peer mode, event kernel)
    uint32_t arr_epoch = 0xffffffffu;           // cycle (vault epoch) the arrival region was last emptied for
    bool peer_multi_domain = false;             // some rank owns more than one domain (qsb_peer_connect)
    const qsb_base_particle* host_in = nullptr;
    unsigned long long n_in_aos = 0;            // tickets [0, n_in_aos) are streamed host records this cycle
    qsb_base_particle* host_out = nullptr;
    unsigned long long host_out_cap = 0, census_copied = 0;   // records already on their way to host_out
    size_t next_chunk = 0;
    unsigned chunk_shift = 18;                  // log2(records per streaming chunk), fixed at creation
    // peer exchange over NVLink (qsb_peer_export / qsb_peer_connect, see include/qsb.h)
    char* vault_base[2] = { nullptr, nullptr }; // the vaults' allocations (header + SoA arrays)
    char* peer_block = nullptr;                 // own exported allocation: vault_base[0] (PeerControl in its header)
    char* peer_base[kMaxPeers] = { nullptr };   // every rank's block as mapped into this process (own rank: peer_block)
    bool peer_on = false;
    uint32_t peer_epoch = 0;
    unsigned long long watchdog_ns = 0;
    PeerControl* h_peer = nullptr;              // pinned: values on their way to / back from the own control block
    // cycleInit on the device (qsb_cycle_init_resident, cycle_init_kernels.cu)
    const double* d_cell_volume = nullptr;               // [n_cells]
    const unsigned long long* d_cell_id = nullptr;       // [n_cells]
    int* d_source_offsets = nullptr;                     // [n_cells+1] the caller's source plan, kept while its plan_id stands
    unsigned long long* d_source_tally = nullptr;        // [n_cells] running source counts, advanced on the device
    bool have_plan = false;
    uint64_t plan_id = 0;
    unsigned long long plan_n_source = 0;
    unsigned long long* d_spectrum = nullptr;            // [n_groups+1] scratch of qsb_census_energy_spectrum
    CycleInitCounters* d_init = nullptr;
    CycleInitCounters* h_init = nullptr;                 // pinned
    uint32_t epoch = 0;
    uint64_t launches = 0;
    int grid = 0, block = 128, regs = 0, blocks_per_sm = 0;
    bool event_mode = false;                    // tracking_mode bit 0: the event-based kernel (track_event_kernels.cu)
    int evt_smem = 0, evt_slots = 0;
    bool in_cycle = false;
    std::string error;
};

namespace {

// one allocation per vault: a header (the PeerControl block of the peer exchange) followed by the SoA arrays, so that the
// whole processing vault can be exported to the other ranks with a single CUDA IPC handle
char* allocVault(qsb_ctx* c, VaultView& v, unsigned long long cap)
{
    char* base = devAlloc<char>(vault_bytes(cap), c->owned);
    QSB_CUDA(cudaMemset(base, 0, kVaultHeaderBytes));
    v = vault_view(base, cap);
    QSB_CUDA(cudaMemset(v.ready, 0, cap * sizeof(uint32_t)));
    return base;
}

void pushControl(qsb_ctx* c)
{
    QSB_CUDA(cudaMemcpyAsync(c->d_ctl, c->h_ctl, sizeof(DevControl), cudaMemcpyHostToDevice, c->stream));
}

void pullControl(qsb_ctx* c)
{
    QSB_CUDA(cudaMemcpyAsync(c->h_ctl, c->d_ctl, sizeof(DevControl), cudaMemcpyDeviceToHost, c->stream));
    QSB_CUDA(cudaStreamSynchronize(c->stream));
}

template <typename F>
int guarded(qsb_ctx* c, F&& body)
{
    if (!c) return QSB_ERR_ARG;
    try
    {
        QSB_CUDA(cudaSetDevice(c->device));
        return body();
    }
    catch (const CudaFailure& f) { c->error = f.what; return QSB_ERR_CUDA; }
    catch (const std::exception& e) { c->error = e.what(); return QSB_ERR_INTERNAL; }
}

thread_local std::string g_create_error;

unsigned chunkShiftFromEnv()
{
    const char* e = std::getenv("QSB_STREAM_CHUNK_LOG2");
    const int v = e ? std::atoi(e) : 18;
    return (unsigned)std::min(std::max(v, 10), 24);
}

} // namespace

extern "C" {

const char* qsb_last_error(qsb_ctx* c) { return c ? c->error.c_str() : g_create_error.c_str(); }
uint64_t qsb_launch_count(qsb_ctx* c) { return c ? c->launches : 0; }
#ifndef QSB_KERNEL_HASH
#define QSB_KERNEL_HASH "unknown"
#endif
const char* qsb_kernel_hash(void) { return QSB_KERNEL_HASH; }
uint64_t qsb_exchange_record_bytes(void) { return sizeof(ExchangeRecord); }

int qsb_create(int device, const qsb_image* image, double time_step, const qsb_options* opt, qsb_ctx** out)
{
    if (!image || !out || image->abi_version != QSB_ABI_VERSION) return QSB_ERR_ARG;
    static_assert(sizeof(ExchangeRecord) == 160, "exchange record layout");
    *out = nullptr;
    qsb_ctx* c = new qsb_ctx;
    try
    {
        int n_dev = 0;
        cudaError_t e = cudaGetDeviceCount(&n_dev);
        if (e != cudaSuccess || n_dev == 0)
            throw CudaFailure{ std::string("no CUDA device usable (") + cudaGetErrorString(e) + "); this library has no CPU path" };
        if (device < 0 || device >= n_dev) throw CudaFailure{ "device index out of range" };
        QSB_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        QSB_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            throw CudaFailure{ std::string("kernels are built for sm_100a only; device is ") + prop.name };
        c->device = device;
        c->sm_count = prop.multiProcessorCount;
        if (opt) c->opt = *opt; else { c->opt = qsb_options{}; c->opt.validation = 1; }
        c->dt = time_step;
        c->chunk_shift = chunkShiftFromEnv();
        c->n_ranks = image->n_ranks; c->my_rank = image->my_rank;
        if (c->n_ranks > 64) throw CudaFailure{ "at most 64 ranks supported by the control block" };
        QSB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        QSB_CUDA(cudaEventCreate(&c->ev0));
        QSB_CUDA(cudaEventCreate(&c->ev1));
        QSB_CUDA(cudaStreamCreateWithFlags(&c->stream_in, cudaStreamNonBlocking));
        QSB_CUDA(cudaStreamCreateWithFlags(&c->stream_out, cudaStreamNonBlocking));
        QSB_CUDA(cudaEventCreateWithFlags(&c->ev_ctl, cudaEventDisableTiming));
        QSB_CUDA(cudaEventCreate(&c->ev_in_done));
        QSB_CUDA(cudaEventCreateWithFlags(&c->ev_stage[0], cudaEventDisableTiming));
        QSB_CUDA(cudaEventCreateWithFlags(&c->ev_stage[1], cudaEventDisableTiming));

        // ---- image ----
        // a fission emits (int)(nuBar + r) neutrons, r in [0, 1): the kernel holds at most 4 (energy, angle) pairs per collision
        // (the reference's MAX_PRODUCTION_SIZE assert, src/CollisionEvent.cc:95-97, is 4 as well)
        for (int m = 0; m < image->n_materials; ++m)
            if (!(image->mat_nu_bar[m] < 4.0)) throw CudaFailure{ "image: a material's nuBar is 4 or more; at most 4 neutrons per fission are supported" };
        const size_t nc = image->n_cells, ng = image->n_groups, nm = image->n_materials, mr = image->max_reactions_per_material;
        DevImage& im = c->im;
        im.n_cells = image->n_cells; im.n_groups = image->n_groups; im.n_materials = image->n_materials;
        im.max_react = image->max_reactions_per_material; im.n_domains = image->n_domains; im.n_ranks = image->n_ranks;
        im.planes = reinterpret_cast<const double4*>(devUpload(image->planes, nc * 96, c->owned));
        im.nodes = devUpload(image->nodes, nc * 42, c->owned);
        im.face_adj_cell = devUpload(image->face_adj_cell, nc * 6, c->owned);
        im.face_event = devUpload(image->face_event, nc * 6, c->owned);
        im.face_adj_domain = devUpload(image->face_adj_domain, nc * 6, c->owned);
        im.face_nbr_rank = devUpload(image->face_nbr_rank, nc * 6, c->owned);
        im.cell_material = devUpload(image->cell_material, nc, c->owned);
        im.domain_cell_offset = devUpload(image->domain_cell_offset, (size_t)image->n_domains + 1, c->owned);
        c->d_domain_offset = im.domain_cell_offset;
        if (image->cell_volume) c->d_cell_volume = devUpload(image->cell_volume, nc, c->owned);
        if (image->cell_id)     c->d_cell_id = reinterpret_cast<const unsigned long long*>(devUpload(image->cell_id, nc, c->owned));
        c->host_domain_offset.assign(image->domain_cell_offset, image->domain_cell_offset + image->n_domains + 1);

        // hot block: energies | xs_total | xs_react | material records, one allocation, 256-byte sub-alignment
        auto pad = [](size_t b) { return (b + 255) & ~size_t(255); };
        const size_t b_energy = pad((ng + 1) * 8), b_total = pad(nm * ng * 8), b_react = pad(nm * ng * mr * 8);
        const size_t b_mass = pad(nm * 8), b_nubar = pad(nm * 8), b_niso = pad(nm * 4), b_nreact = pad(nm * 4);
        const size_t b_rtype = pad(nm * mr), b_periodic = pad(nm);
        const size_t hot_bytes = b_energy + b_total + b_react + b_mass + b_nubar + b_niso + b_nreact + b_rtype + b_periodic;
        char* hot = devAlloc<char>(hot_bytes, c->owned);
        std::vector<char> h(hot_bytes, 0);
        size_t o = 0;
        auto place = [&](const void* src, size_t bytes, size_t padded) { std::memcpy(h.data() + o, src, bytes); char* p = hot + o; o += padded; return p; };
        im.energies = (const double*)place(image->energies, (ng + 1) * 8, b_energy);
        im.xs_total = (const double*)place(image->xs_total, nm * ng * 8, b_total);
        im.xs_react = (const double*)place(image->xs_react, nm * ng * mr * 8, b_react);
        std::vector<double> inv_mass(nm);
        for (size_t m = 0; m < nm; ++m) inv_mass[m] = 1.0 / image->mat_mass[m];
        im.mat_inv_mass = (const double*)place(inv_mass.data(), nm * 8, b_mass);
        im.mat_nu_bar = (const double*)place(image->mat_nu_bar, nm * 8, b_nubar);
        im.mat_n_iso = (const int*)place(image->mat_n_isotopes, nm * 4, b_niso);
        im.mat_n_react = (const int*)place(image->mat_n_reactions, nm * 4, b_nreact);
        im.mat_react_type = (const uint8_t*)place(image->mat_react_type, nm * mr, b_rtype);
        im.mat_periodic = (const uint8_t*)place(image->mat_periodic, nm, b_periodic);
        QSB_CUDA(cudaMemcpy(hot, h.data(), hot_bytes, cudaMemcpyHostToDevice));

        // ---- compact cell records + {total, 1/total} pairs ----
        {
            const int nx = image->global_nx, ny = image->global_ny;
            im.dx = image->global_lx / image->global_nx;      // the host grid's own expressions (src/GlobalFccGrid.cc:26-28)
            im.dy = image->global_ly / image->global_ny;
            im.dz = image->global_lz / image->global_nz;
            im.margin = 1e-6 * std::min(im.dx, std::min(im.dy, im.dz));
            {
                // energy-group estimate (track_kernels.cu: energy_group): edges are e[i] ~ e[0] * (e[n-1]/e[0])^(i/(n)) up to
                // the reference's spacing quirk; only an estimate, the kernel corrects it against the table
                const double lo = std::log2(std::max(image->energies[0], 1e-300));
                const double second = ng >= 2 ? std::log2(std::max(image->energies[1], 1e-300)) : lo + 1.0;
                const double step = second > lo ? second - lo : 1.0;
                im.group_log2_lo = (float)lo;
                im.group_inv_dlog2 = (float)(1.0 / step);
            }
            im.inv_hx = 2.0 / im.dx; im.inv_hy = 2.0 / im.dy; im.inv_hz = 2.0 / im.dz;
            std::vector<CellRec> recs(nc);
            bool compact = image->global_nx < 65535 && image->global_ny < 65535 && image->global_nz < 65535 && nm <= 255;
            const double cell_size[3] = { im.dx, im.dy, im.dz };
            for (size_t cidx = 0; cidx < nc; ++cidx)
            {
                CellRec& r = recs[cidx];
                std::memset(&r, 0, sizeof(r));
                const int gid = image->cell_gid[cidx];
                const int ijk[3] = { gid % nx, (gid / nx) % ny, gid / (nx * ny) };
                r.ix = (uint16_t)ijk[0]; r.iy = (uint16_t)ijk[1]; r.iz = (uint16_t)ijk[2];
                r.material = (uint8_t)image->cell_material[cidx];
                for (int face = 0; face < 6; ++face)
                {
                    r.events |= (uint32_t)(image->face_event[cidx * 6 + face] & 0xf) << (4 * face);
                    r.adj[face] = image->face_adj_cell[cidx * 6 + face];
                }
                // corner nodes must be index * cell size exactly (points 0 and 7 of the 14-point list)
                const double* nodes = image->nodes + cidx * 42;
                for (int ax = 0; ax < 3; ++ax)
                    if (nodes[ax] != ijk[ax] * cell_size[ax] || nodes[21 + ax] != (ijk[ax] + 1) * cell_size[ax]) compact = false;
                for (int f = 0; f < 24 && compact; ++f)
                {
                    const double* pl = image->planes + (cidx * 24 + f) * 4;
                    const int face = f / 4, ax = face / 2;
                    const double sign = (face & 1) ? -1.0 : 1.0;
                    const double coord = (face & 1) ? ijk[ax] * cell_size[ax] : (ijk[ax] + 1) * cell_size[ax];
                    unsigned code = 0;
                    const double below_one = 0.99999999999999988897769753748;   // 1 - 2^-53
                    if (pl[ax] == sign * 1.0) code = 0;
                    else if (pl[ax] == sign * below_one) code = 1;
                    else compact = false;
                    for (int o = 0; o < 3; ++o) if (o != ax && pl[o] != 0.0) compact = false;
                    long long cb, db;
                    const double dabs = std::fabs(pl[3]);
                    std::memcpy(&cb, &coord, 8); std::memcpy(&db, &dabs, 8);
                    const long long k = db - cb;
                    if (k == 0) code |= 0; else if (k == 1) code |= 2; else if (k == -1) code |= 4; else compact = false;
                    // sign of D: -sign * |D| (either sign of zero is acceptable, the fast path never meets it)
                    if (dabs != 0.0 && ((pl[3] < 0) != (sign > 0))) compact = false;
                    r.code[f] = (uint8_t)code;
                }
            }
            im.compact = compact ? 1 : 0;
            if (!compact && !c->opt.validation)
                throw CudaFailure{ "image: the fast kernels need the uniform brick mesh of GlobalFccGrid (axis-aligned facet planes, < 65535 cells per axis, "
                                   "<= 255 materials); this image is not one: use qsb_options.validation = 1" };
            im.cells = devUpload(recs.data(), nc, c->owned);
            // computed-neighbour path (DevImage::brick): the three strides are read off the first on-processor transit of each
            // axis and then checked on EVERY transit face of every cell; any exception (several domains per rank, a domain that
            // is not one lexicographically numbered brick) leaves the kernels on the adjacency records
            {
                int stride[3] = { 0, 0, 0 };
                bool ok = nm <= 255;
                for (size_t cidx = 0; cidx < nc && ok; ++cidx)
                    for (int face = 0; face < 6; ++face)
                    {
                        if ((image->face_event[cidx * 6 + face] & 0xf) != QSB_ADJ_TRANSIT_ON) continue;
                        const long long d = (long long)image->face_adj_cell[cidx * 6 + face] - (long long)cidx;
                        const long long want = (face & 1) ? -d : d;            // faces 0,2,4 look towards +x,+y,+z, faces 1,3,5 towards -x,-y,-z
                        const int ax = face >> 1;
                        if (stride[ax] == 0 && want > 0 && want < (1ll << 30)) stride[ax] = (int)want;
                        if (want != stride[ax] || want <= 0) { ok = false; break; }
                        // the neighbour's grid position must be this cell's, one step along the axis (the kernel derives it)
                        const int g0 = image->cell_gid[cidx], g1 = image->cell_gid[image->face_adj_cell[cidx * 6 + face]];
                        const int step = ax == 0 ? 1 : (ax == 1 ? nx : nx * ny);
                        if (g1 - g0 != ((face & 1) ? -step : step)) { ok = false; break; }
                    }
                std::vector<uint32_t> info(nc);
                for (size_t cidx = 0; cidx < nc; ++cidx) info[cidx] = (recs[cidx].events & 0xffffffu) | ((uint32_t)recs[cidx].material << 24);
                im.cell_info = devUpload(info.data(), nc, c->owned);
                // cells within `depth` cells of a face that leads to another rank (DevImage::cell_near): breadth-first over the
                // on-rank adjacency, starting from the cells that have such a face
                im.cell_near = nullptr;
                if (image->n_ranks > 1)
                {
                    int depth = 8;
                    if (const char* e = std::getenv("QSB_BOUNDARY_DEPTH")) depth = std::max(0, std::atoi(e));
                    std::vector<uint8_t> near(nc, 0);
                    std::vector<uint32_t> frontier, next;
                    for (size_t cidx = 0; cidx < nc; ++cidx)
                        for (int face = 0; face < 6; ++face)
                            if (image->face_event[cidx * 6 + face] == QSB_ADJ_TRANSIT_OFF) { near[cidx] = 1; frontier.push_back((uint32_t)cidx); break; }
                    for (int d = 1; d < depth && !frontier.empty(); ++d)
                    {
                        next.clear();
                        for (uint32_t cidx : frontier)
                            for (int face = 0; face < 6; ++face)
                                if (image->face_event[(size_t)cidx * 6 + face] == QSB_ADJ_TRANSIT_ON)
                                {
                                    const int32_t n = image->face_adj_cell[(size_t)cidx * 6 + face];
                                    if (n >= 0 && (size_t)n < nc && !near[n]) { near[n] = 1; next.push_back((uint32_t)n); }
                                }
                        frontier.swap(next);
                    }
                    if (depth > 0) im.cell_near = devUpload(near.data(), nc, c->owned);
                }
                im.brick = (ok && std::getenv("QSB_NO_BRICK") == nullptr) ? 1 : 0;
                for (int k = 0; k < 3; ++k) im.brick_stride[k] = stride[k];
            }
            // compact reaction table: isotope 0's rows only, when every material's isotopes share one table
            {
                bool all_periodic = true;
                int nr = 0;
                for (size_t m = 0; m < nm; ++m)
                {
                    if (!image->mat_periodic[m] || image->mat_n_reactions[m] > 9) all_periodic = false;
                    nr = std::max(nr, (int)image->mat_n_reactions[m]);
                }
                im.xs_compact = nullptr; im.compact_react = 0;
                if (all_periodic && nr > 0 && std::getenv("QSB_NO_COMPACT_XS") == nullptr)
                {
                    std::vector<double> compactTable(nm * ng * (size_t)nr, 0.0);
                    for (size_t m = 0; m < nm; ++m)
                        for (size_t g = 0; g < ng; ++g)
                            for (int k = 0; k < image->mat_n_reactions[m]; ++k)
                                compactTable[(m * ng + g) * nr + k] = image->xs_react[(m * ng + g) * mr + k];
                    im.xs_compact = devUpload(compactTable.data(), compactTable.size(), c->owned);
                    im.compact_react = nr;
                }
            }
            std::vector<double2> pairs(nm * ng);
            for (size_t i = 0; i < nm * ng; ++i) { pairs[i].x = image->xs_total[i]; pairs[i].y = 1.0 / image->xs_total[i]; }
            im.xs_pair = devUpload(pairs.data(), nm * ng, c->owned);
        }

        // pin the hot block in L2 (126 MB on B200): persisting carve-out + access-policy window on our stream
        {
            size_t want = std::min<size_t>(hot_bytes, (size_t)prop.persistingL2CacheMaxSize);
            // QSB_NO_L2_WINDOW=1: leave the hot block to the ordinary L2 replacement policy (bench.py's NonFlatXC A/B)
            if (want > 0 && prop.accessPolicyMaxWindowSize > 0 && std::getenv("QSB_NO_L2_WINDOW") == nullptr)
            {
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
                cudaStreamAttrValue attr{};
                attr.accessPolicyWindow.base_ptr = hot;
                attr.accessPolicyWindow.num_bytes = std::min<size_t>(hot_bytes, (size_t)prop.accessPolicyMaxWindowSize);
                attr.accessPolicyWindow.hitRatio = 1.0f;
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
                cudaGetLastError();     // the window is a hint; failure to set it is not an error
            }
        }

        // ---- vaults, slabs, tallies ----
        unsigned long long cap = c->opt.particle_capacity;
        if (cap == 0) cap = 1ull << 20;
        cap = (cap + 31ull) & ~31ull;
        c->vault_base[0] = allocVault(c, c->vault[0], cap);
        c->vault_base[1] = allocVault(c, c->vault[1], cap);
        c->send_capacity = c->n_ranks > 1 ? (c->opt.send_capacity ? c->opt.send_capacity : std::max<unsigned long long>(cap / 8, 4096)) : 1;
        c->sends = devAlloc<ExchangeRecord>((size_t)c->n_ranks * c->send_capacity, c->owned);
        c->flux = devAlloc<double>(nc * ng, c->owned);
        QSB_CUDA(cudaMemset(c->flux, 0, nc * ng * sizeof(double)));
        c->flux_partial = devAlloc<double>(1024, c->owned);
        c->d_ctl = devAlloc<DevControl>(1, c->owned);
        QSB_CUDA(cudaMallocHost((void**)&c->h_ctl, sizeof(DevControl)));
        QSB_CUDA(cudaMallocHost((void**)&c->h_partial, 1024 * sizeof(double)));
        std::memset(c->h_ctl, 0, sizeof(DevControl));
        c->staging_records = 1u << 20;
        c->staging = devAlloc<char>(c->staging_records * sizeof(qsb_base_particle), c->owned);
        QSB_CUDA(cudaMallocHost(&c->h_staging, c->staging_records * sizeof(qsb_base_particle)));

        // ---- launch shape: persistent grid, resident blocks per SM from the occupancy calculator ----
        c->block = c->opt.threads_per_block > 0 ? c->opt.threads_per_block : 128;
        // the kernel is compiled with __launch_bounds__(128) and uses full-mask warp intrinsics throughout
        if (c->block < 32 || c->block > 128 || (c->block & 31) != 0)
            throw CudaFailure{ "qsb_options.threads_per_block must be a multiple of 32 between 32 and 128" };
        // which tracking kernel: tracking_mode bit 0 (0: the default, event-based -- measured 15-20 % faster, DESIGN.md
        // section 5; 1: history-based); QSB_TRACKING=history|event overrides (A/B measurements: bench.py)
        c->event_mode = (c->opt.tracking_mode & 1) == 0;
        if (const char* e = std::getenv("QSB_TRACKING"))
        {
            if (std::strcmp(e, "event") == 0) c->event_mode = true;
            else if (std::strcmp(e, "history") == 0) c->event_mode = false;
        }
        if (c->event_mode)
        {
            if (c->opt.validation) track_event_kernel_attributes_validation(&c->regs, &c->blocks_per_sm, &c->block, &c->evt_smem, &c->evt_slots);
            else                   track_event_kernel_attributes_fast(&c->regs, &c->blocks_per_sm, &c->block, &c->evt_smem, &c->evt_slots);
        }
        else if (c->opt.validation) track_kernel_attributes_validation(&c->regs, &c->blocks_per_sm, c->block);
        else                        track_kernel_attributes_fast(&c->regs, &c->blocks_per_sm, c->block);
        check(cudaGetLastError(), "kernel attributes (is the sm_100a image loadable on this device?)");
        if (c->blocks_per_sm < 1) throw CudaFailure{ "tracking kernel cannot be resident on this device" };
        if (c->opt.blocks_per_sm > 0) c->blocks_per_sm = std::min(c->blocks_per_sm, c->opt.blocks_per_sm);
        if (const char* e = std::getenv("QSB_BLOCKS_PER_SM"))         // occupancy experiments only
            if (std::atoi(e) > 0) c->blocks_per_sm = std::min(c->blocks_per_sm, std::atoi(e));
        c->grid = c->sm_count * c->blocks_per_sm;
        if (std::getenv("QSB_TRACE"))
            std::fprintf(stderr, "[qsb] rank %d: %s kernel (%s build), %d registers, %d threads x %d blocks per SM, grid %d, shared memory %d B per block; "
                         "compact geometry %d, brick neighbours %d (strides %d %d %d), compact reaction table %d\n",
                         c->my_rank, c->event_mode ? "event-based" : "history-based", c->opt.validation ? "validation" : "fast", c->regs, c->block,
                         c->blocks_per_sm, c->grid, c->evt_smem, im.compact, im.brick, im.brick_stride[0], im.brick_stride[1], im.brick_stride[2],
                         im.xs_compact ? im.compact_react : 0);
        QSB_CUDA(cudaStreamSynchronize(c->stream));
    }
    catch (const CudaFailure& f)
    {
        g_create_error = f.what;
        const bool bad_arg = f.what.rfind("qsb_options.", 0) == 0 || f.what.rfind("image:", 0) == 0;
        qsb_destroy(c);             // frees device + page-locked allocations, streams and events created so far
        return bad_arg ? QSB_ERR_ARG : QSB_ERR_CUDA;
    }
    *out = c;
    return QSB_OK;
}

int qsb_destroy(qsb_ctx* c)
{
    if (!c) return QSB_ERR_ARG;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    qsb_peer_disconnect(c);
    if (c->h_peer) cudaFreeHost(c->h_peer);
    for (void* p : c->owned) cudaFree(p);
    if (c->h_ctl) cudaFreeHost(c->h_ctl);
    if (c->h_partial) cudaFreeHost(c->h_partial);
    if (c->h_init) cudaFreeHost(c->h_init);
    if (c->h_staging) cudaFreeHost(c->h_staging);
    if (c->h_chunk_flags) cudaFreeHost(c->h_chunk_flags);
    if (c->h_marks) cudaFreeHost(c->h_marks);
    if (c->stream_in) cudaStreamDestroy(c->stream_in);
    if (c->stream_out) cudaStreamDestroy(c->stream_out);
    if (c->ev_ctl) cudaEventDestroy(c->ev_ctl);
    if (c->ev_in_done) cudaEventDestroy(c->ev_in_done);
    for (cudaEvent_t e : c->ev_stage) if (e) cudaEventDestroy(e);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return QSB_OK;
}

int qsb_cycle_begin(qsb_ctx* c, int keep_census)
{
    return guarded(c, [&]() {
        c->epoch++;
        unsigned long long carried = 0;
        if (keep_census && c->in_cycle)
        {
            if (c->streaming)
            { c->error = "keep_census: the census of a streamed cycle (qsb_stream_begin) was delivered to the host, not kept in the vault"; return (int)QSB_ERR_STATE; }
            pullControl(c);
            carried = std::min<unsigned long long>(c->h_ctl->census_count, c->vault[1 - c->proc].capacity);
            c->proc = 1 - c->proc;                           // last cycle's census becomes the processing vault
            // its slots carry an older epoch: they are covered by ready_prefix instead
        }
        std::memset(c->h_ctl, 0, sizeof(DevControl));
        c->h_ctl->epoch = c->epoch;
        c->h_ctl->tail = carried;
        c->host_tail = carried;
        c->ready_prefix = carried;
        c->consumed = 0;
        c->pending_inflight = carried;
        c->streaming = false; c->stream_input_issued = false;
        c->n_in_aos = 0; c->host_in = nullptr; c->host_out = nullptr; c->host_out_cap = 0; c->census_copied = 0; c->next_chunk = 0;
        pushControl(c);
        QSB_CUDA(cudaMemsetAsync(c->flux, 0, (size_t)c->im.n_cells * c->im.n_groups * sizeof(double), c->stream));
        c->in_cycle = true;
        return (int)QSB_OK;
    });
}

int qsb_put_particles(qsb_ctx* c, const qsb_base_particle* aos, uint64_t n)
{
    if (n && !aos) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        if (!c->in_cycle) { c->error = "qsb_put_particles before qsb_cycle_begin"; return (int)QSB_ERR_STATE; }
        if (c->streaming) { c->error = "qsb_put_particles after qsb_stream_begin: the cycle's input is the streamed host vault"; return (int)QSB_ERR_STATE; }
        VaultView& v = c->vault[c->proc];
        if (c->host_tail != c->h_ctl->tail)
        { c->error = "qsb_put_particles after tracking started: use qsb_put_arrivals"; return (int)QSB_ERR_STATE; }
        if (c->host_tail + n > v.capacity)
        { c->error = "processing vault capacity exceeded by qsb_put_particles"; return (int)QSB_ERR_CAPACITY; }
        // chunked: pageable host -> pinned -> device AoS staging -> SoA scatter kernel
        const size_t chunk = c->staging_records;
        for (uint64_t done = 0; done < n; done += chunk)
        {
            const uint64_t m = std::min<uint64_t>(chunk, n - done);
            QSB_CUDA(cudaStreamSynchronize(c->stream));      // pinned staging buffer is free again
            std::memcpy(c->h_staging, aos + done, m * sizeof(qsb_base_particle));
            QSB_CUDA(cudaMemcpyAsync(c->staging, c->h_staging, m * sizeof(qsb_base_particle), cudaMemcpyHostToDevice, c->stream));
            const int grid = (int)std::min<uint64_t>((m + 255) / 256, 4096);
            aos_to_soa_kernel<<<grid, 256, 0, c->stream>>>((const qsb_base_particle*)c->staging, m, v, c->host_tail + done,
                                                           c->d_domain_offset, c->epoch);
            c->launches++;
        }
        QSB_CUDA(cudaGetLastError());
        c->host_tail += n;
        c->ready_prefix = c->host_tail;
        c->h_ctl->tail = c->host_tail;
        c->pending_inflight += n;
        QSB_CUDA(cudaMemcpyAsync(&c->d_ctl->tail, &c->h_ctl->tail, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
        return (int)QSB_OK;
    });
}

// host AoS records -> SoA slots [first, first+n) of vault `v`, staged through the pinned bounce buffer
static void uploadRecords(qsb_ctx* c, VaultView& v, unsigned long long first, const qsb_base_particle* aos, uint64_t n)
{
    const size_t chunk = c->staging_records;
    for (uint64_t done = 0; done < n; done += chunk)
    {
        const uint64_t m = std::min<uint64_t>(chunk, n - done);
        QSB_CUDA(cudaStreamSynchronize(c->stream));      // pinned staging buffer is free again
        std::memcpy(c->h_staging, aos + done, m * sizeof(qsb_base_particle));
        QSB_CUDA(cudaMemcpyAsync(c->staging, c->h_staging, m * sizeof(qsb_base_particle), cudaMemcpyHostToDevice, c->stream));
        const int grid = (int)std::min<uint64_t>((m + 255) / 256, 4096);
        aos_to_soa_kernel<<<grid, 256, 0, c->stream>>>((const qsb_base_particle*)c->staging, m, v, first + done, c->d_domain_offset, c->epoch);
        c->launches++;
    }
    QSB_CUDA(cudaGetLastError());
}

int qsb_put_census(qsb_ctx* c, const qsb_base_particle* aos, uint64_t n)
{
    if (n && !aos) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        VaultView& v = c->vault[1 - c->proc];
        if (n > v.capacity) { c->error = "census vault capacity exceeded by qsb_put_census"; return (int)QSB_ERR_CAPACITY; }
        uploadRecords(c, v, 0, aos, n);
        // from here on the context looks as if a (non-streamed) cycle had just ended with this census
        std::memset(c->h_ctl, 0, sizeof(DevControl));
        c->h_ctl->epoch = c->epoch;
        c->h_ctl->census_count = n;
        pushControl(c);
        QSB_CUDA(cudaStreamSynchronize(c->stream));
        c->host_tail = c->ready_prefix = c->consumed = c->pending_inflight = 0;
        c->n_in_aos = 0;
        c->streaming = false;
        c->in_cycle = true;
        return (int)QSB_OK;
    });
}

int qsb_cycle_init_resident(qsb_ctx* c, const qsb_cycle_init_args* args, qsb_cycle_init_result* result)
{
    if (!args) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        const int n_cells = c->im.n_cells;
        if (!c->d_cell_volume || !c->d_cell_id)
        { c->error = "qsb_cycle_init_resident: the image carries no cell_volume / cell_id arrays"; return (int)QSB_ERR_STATE; }
        if (!(args->source_weight > 0.0) || !(args->split_factor > 0.0))
        { c->error = "qsb_cycle_init_resident: source_weight and split_factor must be positive"; return (int)QSB_ERR_ARG; }
        // ---- last cycle's census: how many records, and that they are in the vault ----
        unsigned long long carried = 0;
        if (c->in_cycle)
        {
            if (c->streaming)
            { c->error = "qsb_cycle_init_resident: the previous cycle's census was streamed to the host (qsb_stream_begin), not kept in the vault; "
                         "hand it back with qsb_put_census"; return (int)QSB_ERR_STATE; }
            pullControl(c);
            carried = std::min<unsigned long long>(c->h_ctl->census_count, c->vault[1 - c->proc].capacity);
        }
        // ---- source plan ----
        if (!c->have_plan || args->plan_id != c->plan_id)
        {
            if (!args->source_offsets || !args->source_tally)
            { c->error = "qsb_cycle_init_resident: a new plan_id needs source_offsets and source_tally"; return (int)QSB_ERR_ARG; }
            if (args->source_offsets[0] != 0) { c->error = "qsb_cycle_init_resident: source_offsets[0] must be 0"; return (int)QSB_ERR_ARG; }
            for (int i = 0; i < n_cells; ++i)
                if (args->source_offsets[i + 1] < args->source_offsets[i])
                { c->error = "qsb_cycle_init_resident: source_offsets must not decrease"; return (int)QSB_ERR_ARG; }
            if (!c->d_source_offsets)
            {
                c->d_source_offsets = devAlloc<int>((size_t)n_cells + 1, c->owned);
                c->d_source_tally = devAlloc<unsigned long long>((size_t)n_cells, c->owned);
                c->d_init = devAlloc<CycleInitCounters>(1, c->owned);
                QSB_CUDA(cudaMallocHost((void**)&c->h_init, sizeof(CycleInitCounters)));
            }
            QSB_CUDA(cudaStreamSynchronize(c->stream));
            QSB_CUDA(cudaMemcpy(c->d_source_offsets, args->source_offsets, ((size_t)n_cells + 1) * sizeof(int), cudaMemcpyHostToDevice));
            QSB_CUDA(cudaMemcpy(c->d_source_tally, args->source_tally, (size_t)n_cells * sizeof(unsigned long long), cudaMemcpyHostToDevice));
            c->plan_n_source = (unsigned long long)args->source_offsets[n_cells];
            c->plan_id = args->plan_id;
            c->have_plan = true;
        }
        // ---- the cycle's clean slate (what qsb_cycle_begin does) ----
        c->epoch++;
        std::memset(c->h_ctl, 0, sizeof(DevControl));
        c->h_ctl->epoch = c->epoch;
        c->consumed = 0;
        c->streaming = false; c->stream_input_issued = false;
        c->n_in_aos = 0; c->host_in = nullptr; c->host_out = nullptr; c->host_out_cap = 0; c->census_copied = 0; c->next_chunk = 0;
        pushControl(c);
        QSB_CUDA(cudaMemsetAsync(c->flux, 0, (size_t)c->im.n_cells * c->im.n_groups * sizeof(double), c->stream));
        QSB_CUDA(cudaMemsetAsync(c->d_init, 0, sizeof(CycleInitCounters), c->stream));
        // ---- census + source -> population control -> roulette -> processing vault ----
        CycleInitArgs a;
        a.src = c->vault[1 - c->proc];
        a.dst = c->vault[c->proc];
        a.n_carried = carried; a.n_source = c->plan_n_source;
        a.source_offsets = c->d_source_offsets; a.source_tally = c->d_source_tally;
        a.cell_id = c->d_cell_id; a.cell_volume = c->d_cell_volume; a.nodes = c->im.nodes; a.n_cells = n_cells;
        a.source_weight = args->source_weight; a.e_min = args->e_min; a.e_max = args->e_max; a.dt = c->dt;
        a.factor = args->split_factor;
        a.cutoff = args->low_weight_cutoff; a.weight_cutoff = args->low_weight_cutoff * args->source_weight;
        a.epoch = c->epoch;
        a.out = c->d_init;
        uint32_t n_launch = 0;
        QSB_CUDA(cudaEventRecord(c->ev0, c->stream));
        if (a.n_carried + a.n_source > 0)
        {
            launch_cycle_init(a, c->sm_count, c->stream);
            ++n_launch;
            if (a.n_source > 0) { launch_source_tally_advance(c->d_source_tally, c->d_source_offsets, n_cells, c->stream); ++n_launch; }
            QSB_CUDA(cudaGetLastError());
            c->launches += n_launch;
        }
        QSB_CUDA(cudaEventRecord(c->ev1, c->stream));
        QSB_CUDA(cudaMemcpyAsync(c->h_init, c->d_init, sizeof(CycleInitCounters), cudaMemcpyDeviceToHost, c->stream));
        QSB_CUDA(cudaStreamSynchronize(c->stream));
        float ms = 0;
        QSB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        c->in_cycle = true;
        const unsigned long long n_out = std::min<unsigned long long>(c->h_init->n_out, a.dst.capacity);
        c->h_ctl->tail = n_out;
        c->host_tail = n_out;
        c->ready_prefix = n_out;
        c->pending_inflight = n_out;
        QSB_CUDA(cudaMemcpyAsync(&c->d_ctl->tail, &c->h_ctl->tail, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
        if (result)
        {
            result->n_start = carried; result->n_source = a.n_source;
            result->n_rr = c->h_init->n_rr; result->n_split = c->h_init->n_split;
            result->n_processing = n_out; result->device_ms = ms; result->n_launches = n_launch;
        }
        if (c->h_init->overflow || c->h_init->n_out > a.dst.capacity)
        { c->error = "processing vault capacity exceeded by qsb_cycle_init_resident; raise qsb_options.particle_capacity"; return (int)QSB_ERR_CAPACITY; }
        return (int)QSB_OK;
    });
}

int qsb_put_arrivals(qsb_ctx* c, const void* device_records, uint64_t n)
{
    if (n && !device_records) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        if (!c->in_cycle) { c->error = "qsb_put_arrivals before qsb_cycle_begin"; return (int)QSB_ERR_STATE; }
        if (n == 0) return (int)QSB_OK;
        pullControl(c);
        VaultView& v = c->vault[c->proc];
        const unsigned long long first = c->h_ctl->tail - c->n_in_aos;      // SoA slot: tickets below n_in_aos are streamed records
        if (first + n > v.capacity) { c->error = "processing vault capacity exceeded by arrivals"; return (int)QSB_ERR_CAPACITY; }
        const int grid = (int)std::min<uint64_t>((n + 255) / 256, 4096);
        arrivals_to_soa_kernel<<<grid, 256, 0, c->stream>>>((const ExchangeRecord*)device_records, n, v, first, c->d_domain_offset, c->epoch);
        c->launches++;
        QSB_CUDA(cudaGetLastError());
        c->h_ctl->tail += n;
        c->pending_inflight += n;
        QSB_CUDA(cudaMemcpyAsync(&c->d_ctl->tail, &c->h_ctl->tail, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
        QSB_CUDA(cudaStreamSynchronize(c->stream));
        return (int)QSB_OK;
    });
}

} // extern "C"

namespace {

// streaming chunk: 2^18 records = 35.7 MB per DMA copy by default.  Every input chunk is followed by an 8-byte copy that
// moves ctl->in_ready (stream order), which costs the copy engine a fixed ~20-30 us; at 8.9 MB chunks that was a fifth of the
// transfer time (measured), at 35.7 MB it is noise, and the pipeline still starts after 0.7 ms.  QSB_STREAM_CHUNK_LOG2 overrides.
#define kCensusChunkShift (c->chunk_shift)
#define kInputChunkRecords ((size_t)1 << c->chunk_shift)

bool isPinnedHost(const void* p)
{
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return attr.type == cudaMemoryTypeHost;
}

// enqueue the H2D copies of the streamed host vault, chunk by chunk, each followed by an 8-byte copy that moves
// ctl->in_ready forward (stream order: a chunk has landed before its mark).  Pinned source: everything is enqueued at
// once and runs on the copy engine while the kernel tracks.  Pageable source: staged through the two halves of the
// pinned bounce buffer by this thread, census chunks serviced in between.
void serviceCensus(qsb_ctx* c, bool final);

void issueStreamInput(qsb_ctx* c)
{
    if (c->stream_input_issued || c->n_in_aos == 0) { c->stream_input_issued = true; return; }
    c->stream_input_issued = true;
    const unsigned long long n = c->n_in_aos;
    // (h_marks was sized by qsb_stream_begin: page-locked allocation / free synchronise implicitly with the running kernel,
    //  which at this point is spinning on the very marks this function is about to enqueue)
    const bool pinned = isPinnedHost(c->host_in);
    const size_t half = c->staging_records / 2;
    size_t k = 0;
    for (unsigned long long done = 0; done < n; ++k)
    {
        const unsigned long long m = std::min<unsigned long long>(pinned ? kInputChunkRecords : std::min(kInputChunkRecords, half), n - done);
        const void* src = c->host_in + done;
        if (!pinned)
        {
            const int h = (int)(k & 1);
            QSB_CUDA(cudaEventSynchronize(c->ev_stage[h]));             // this half's previous copy has left
            char* bounce = (char*)c->h_staging + (size_t)h * half * sizeof(qsb_base_particle);
            std::memcpy(bounce, src, m * sizeof(qsb_base_particle));
            src = bounce;
        }
        QSB_CUDA(cudaMemcpyAsync(c->d_in_aos + done, src, m * sizeof(qsb_base_particle), cudaMemcpyHostToDevice, c->stream_in));
        if (!pinned) QSB_CUDA(cudaEventRecord(c->ev_stage[k & 1], c->stream_in));
        done += m;
        if (k >= c->n_marks) throw CudaFailure{ "internal: input chunk marks exhausted" };
        c->h_marks[k] = done;
        QSB_CUDA(cudaMemcpyAsync(&c->d_ctl->in_ready, &c->h_marks[k], sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream_in));
        if (!pinned) serviceCensus(c, false);
    }
    QSB_CUDA(cudaEventRecord(c->ev_in_done, c->stream_in));
}

// start the D2H copy of every census chunk the kernel has announced (final = false), or of everything that is left
// once the kernel has ended and the control block is back (final = true)
void serviceCensus(qsb_ctx* c, bool final)
{
    if (!c->streaming || !c->host_out) return;
    const unsigned long long chunk = 1ull << kCensusChunkShift;
    while (c->next_chunk < c->n_chunks && *((volatile unsigned int*)&c->h_chunk_flags[c->next_chunk]) == c->epoch)
    {
        const unsigned long long first = (unsigned long long)c->next_chunk * chunk;
        if (first + chunk > c->host_out_cap) break;                       // the caller's buffer ends here; the rest is fetched later
        QSB_CUDA(cudaMemcpyAsync(c->host_out + first, c->d_census_aos + first, chunk * sizeof(qsb_base_particle),
                                 cudaMemcpyDeviceToHost, c->stream_out));
        c->census_copied = first + chunk;
        c->next_chunk++;
    }
    if (final)
    {
        const unsigned long long n = std::min<unsigned long long>(std::min<unsigned long long>(c->h_ctl->census_count, c->vault[0].capacity), c->host_out_cap);
        if (n > c->census_copied)
        {
            QSB_CUDA(cudaMemcpyAsync(c->host_out + c->census_copied, c->d_census_aos + c->census_copied,
                                     (n - c->census_copied) * sizeof(qsb_base_particle), cudaMemcpyDeviceToHost, c->stream_out));
            c->census_copied = n;
            c->next_chunk = (size_t)(n >> kCensusChunkShift);
        }
    }
}

} // namespace

extern "C" {

int qsb_stream_begin(qsb_ctx* c, const qsb_base_particle* in, uint64_t n_in, qsb_base_particle* census_out, uint64_t census_cap)
{
    if ((n_in && !in) || (census_cap && !census_out)) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        if (!c->in_cycle) { c->error = "qsb_stream_begin before qsb_cycle_begin"; return (int)QSB_ERR_STATE; }
        if (c->streaming || c->host_tail != c->ready_prefix || c->h_ctl->tail != c->host_tail)
        { c->error = "qsb_stream_begin must directly follow qsb_cycle_begin"; return (int)QSB_ERR_STATE; }
        const unsigned long long cap = c->vault[0].capacity;
        if (n_in > c->in_aos_cap)
        {
            c->in_aos_cap = std::max<size_t>(n_in + n_in / 4, 1u << 20);
            c->d_in_aos = devAlloc<qsb_base_particle>(c->in_aos_cap, c->owned);      // the previous, smaller one stays owned until destroy
        }
        if (!c->d_census_aos)
        {
            c->d_census_aos = devAlloc<qsb_base_particle>(cap, c->owned);
            c->n_chunks = (size_t)((cap + (1ull << kCensusChunkShift) - 1) >> kCensusChunkShift);
            c->d_chunk_done = devAlloc<unsigned int>(c->n_chunks, c->owned);
            QSB_CUDA(cudaHostAlloc((void**)&c->h_chunk_flags, c->n_chunks * sizeof(unsigned int), cudaHostAllocMapped));
            std::memset(c->h_chunk_flags, 0, c->n_chunks * sizeof(unsigned int));
            QSB_CUDA(cudaHostGetDevicePointer((void**)&c->d_chunk_flags, c->h_chunk_flags, 0));
        }
        {
            // one in_ready mark per input copy; a pageable source is staged through half the bounce buffer per copy.  Sized
            // here, BEFORE the tracking kernel is launched: cudaMallocHost / cudaFreeHost may block until a running kernel ends,
            // and the kernel cannot end before the marks have been enqueued.
            const size_t per_copy = std::min(kInputChunkRecords, std::max<size_t>(c->staging_records / 2, 1));
            const size_t need = (size_t)((n_in + per_copy - 1) / per_copy) + 1;
            if (need > c->n_marks)
            {
                QSB_CUDA(cudaStreamSynchronize(c->stream_in));
                if (c->h_marks) { QSB_CUDA(cudaFreeHost(c->h_marks)); c->h_marks = nullptr; c->n_marks = 0; }
                QSB_CUDA(cudaMallocHost((void**)&c->h_marks, (need + 64) * sizeof(unsigned long long)));
                c->n_marks = need + 64;
            }
        }
        QSB_CUDA(cudaMemsetAsync(c->d_chunk_done, 0, c->n_chunks * sizeof(unsigned int), c->stream));
        c->streaming = true; c->stream_input_issued = false;
        c->host_in = in; c->n_in_aos = n_in;
        c->host_out = census_out; c->host_out_cap = census_cap; c->census_copied = 0; c->next_chunk = 0;
        c->h_ctl->tail += n_in;                       // ticket space: [0, n_in) streamed records, then the SoA slots
        c->host_tail = c->h_ctl->tail;
        c->pending_inflight += n_in;
        QSB_CUDA(cudaMemcpyAsync(&c->d_ctl->tail, &c->h_ctl->tail, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
        return (int)QSB_OK;
    });
}

int qsb_track(qsb_ctx* c, qsb_track_stats* stats)
{
    return guarded(c, [&]() {
        if (!c->in_cycle) { c->error = "qsb_track before qsb_cycle_begin"; return (int)QSB_ERR_STATE; }
        TrackArgs a;
        a.im = c->im;
        a.proc = c->vault[c->proc];
        a.census = c->vault[1 - c->proc];
        a.sends = c->sends; a.send_capacity = c->send_capacity;
        a.ctl = c->d_ctl; a.flux = c->flux; a.dt = c->dt; a.ready_prefix = c->ready_prefix;
        a.epoch = c->epoch;
        a.check_mode = ((c->opt.tracking_mode & 2) ? 1 : 0) | ((c->opt.tracking_mode & 4) ? 2 : 0);
        a.in_aos = c->d_in_aos; a.n_in = c->n_in_aos;
        a.inflight = &c->d_ctl->inflight; a.tail = &c->d_ctl->tail;
        a.peer_mode = 0; a.peer_multi_domain = 0; a.arrival_first = 0; a.arrival_cap = 0; a.prio_list = nullptr; a.prio_slots = 0; a.my_rank = c->my_rank;
        uint32_t n_launch_pre = 0; a.peer_epoch = 0; a.watchdog_ns = c->watchdog_ns;
        for (int r = 0; r < kMaxPeers; ++r) a.peer_base[r] = c->peer_base[r];
        if (c->peer_on)
        {
            if (c->proc != 0) { c->error = "peer exchange: the exported processing vault is vault 0 (keep_census is not supported with it)"; return (int)QSB_ERR_STATE; }
            a.peer_mode = 1;
            a.peer_multi_domain = c->peer_multi_domain ? 1 : 0;
#if QSB_OPT_BOUNDARY_FIRST
            // boundary first: list the slots of the cycle's initial population whose history may reach another GPU (first launch of
            // a cycle over a host-put / device-made population; not when the input is streamed or a launch continues a cycle)
            // OPT-IN (QSB_BOUNDARY_FIRST=1): parity-green and 3 % faster on 2 GPUs (profiles/r02_multi_gpu_lines.txt, call 24), but the
            // round's GPU budget ended before it could be run on 4 and 8
            static const bool boundary_first = [] { const char* e = std::getenv("QSB_BOUNDARY_FIRST"); return e && std::atoi(e) != 0; }();
            if (boundary_first && c->event_mode && c->im.cell_near && c->consumed == 0 && c->n_in_aos == 0 && c->ready_prefix > 0 &&
                c->ready_prefix < (1ull << 32))
            {
                if (c->ready_prefix > c->prio_list_cap)
                {
                    c->prio_list_cap = (size_t)(c->ready_prefix + c->ready_prefix / 4 + 1024);
                    c->d_prio_list = devAlloc<uint32_t>(c->prio_list_cap, c->owned);      // the previous, smaller one stays owned until destroy
                }
                QSB_CUDA(cudaMemsetAsync(&c->d_ctl->prio_head, 0, sizeof(unsigned long long), c->stream));
                QSB_CUDA(cudaMemsetAsync(&c->d_ctl->prio_count, 0, sizeof(unsigned long long), c->stream));
                const int grid = (int)std::min<unsigned long long>((c->ready_prefix + 255) / 256, (unsigned long long)c->sm_count * 16);
                boundary_list_kernel<<<grid, 256, 0, c->stream>>>(a.proc.cell, c->ready_prefix, c->im.cell_near, c->d_prio_list, &c->d_ctl->prio_count);
                QSB_CUDA(cudaGetLastError());
                ++n_launch_pre; c->launches++;
                a.prio_list = c->d_prio_list; a.prio_slots = c->ready_prefix;
            }
#endif
#if QSB_OPT_ARRIVAL_QUEUE
            if (c->event_mode)
            {
                a.arrival_cap = a.proc.capacity / 8;
                a.arrival_first = a.proc.capacity - a.arrival_cap;
                if (c->h_ctl->tail - c->n_in_aos > a.arrival_first)
                { c->error = "peer exchange: the processing vault's population reaches into the arrival region (the last eighth of particle_capacity); raise qsb_options.particle_capacity"; return (int)QSB_ERR_CAPACITY; }
            }
#endif
            a.peer_epoch = ++c->peer_epoch;
            PeerControl* d_peer = reinterpret_cast<PeerControl*>(c->peer_block);
            a.inflight = &d_peer->inflight; a.tail = &d_peer->tail;
        }
        a.census_aos = c->streaming ? c->d_census_aos : nullptr;
        a.census_chunk_done = c->d_chunk_done; a.host_chunk_flags = c->d_chunk_flags; a.census_chunk_shift = kCensusChunkShift;
        uint32_t n_launch = n_launch_pre;
        // tickets handed out past the tail by the previous call were never redeemed: restart at the consumed mark
        c->h_ctl->head = c->consumed;
        if (c->event_mode)
        {
            // two queues: streamed input records [0, n_in) and vault slots (tickets >= n_in), see wq_load
            c->h_ctl->head_in = std::min<unsigned long long>(c->consumed, c->n_in_aos);
            c->h_ctl->head = std::max<unsigned long long>(c->consumed, c->n_in_aos);
            QSB_CUDA(cudaMemcpyAsync(&c->d_ctl->head_in, &c->h_ctl->head_in, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
        }
        c->h_ctl->inflight = c->pending_inflight;
        c->pending_inflight = 0;
        QSB_CUDA(cudaMemcpyAsync(&c->d_ctl->head, &c->h_ctl->head, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
        QSB_CUDA(cudaMemcpyAsync(&c->d_ctl->inflight, &c->h_ctl->inflight, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
        if (c->peer_on)
        {
            // the queue state of this launch, THEN the epoch that tells the other ranks it is in place (senders and the
            // termination waves wait for it)
            PeerControl* d_peer = reinterpret_cast<PeerControl*>(c->peer_block);
            c->h_peer->inflight = c->h_ctl->inflight;
            c->h_peer->tail = c->h_ctl->tail;
            // the arrival region: emptied with the first launch of a cycle (slots are stamped with the cycle's epoch); a later
            // launch of the same cycle goes on where the last one stopped -- slots claimed past the tail by warps that are gone
            // were never redeemed
            if (c->arr_epoch != c->epoch) { c->arr_epoch = c->epoch; c->h_peer->arr_tail = 0; c->h_ctl->arr_head = 0; }
            else c->h_ctl->arr_head = std::min(c->h_ctl->arr_head, c->h_peer->arr_tail);
            QSB_CUDA(cudaMemcpyAsync(&d_peer->arr_tail, &c->h_peer->arr_tail, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
            QSB_CUDA(cudaMemcpyAsync(&c->d_ctl->arr_head, &c->h_ctl->arr_head, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
            c->h_peer->n_in = c->n_in_aos;
            c->h_peer->vault_epoch = c->epoch;
            c->h_peer->epoch = c->peer_epoch;
            QSB_CUDA(cudaMemcpyAsync(&d_peer->inflight, &c->h_peer->inflight, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
            QSB_CUDA(cudaMemcpyAsync(&d_peer->tail, &c->h_peer->tail, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
            QSB_CUDA(cudaMemsetAsync(&d_peer->first_idle_ns, 0, 6 * sizeof(unsigned long long), c->stream));          // this launch's diagnostics
            QSB_CUDA(cudaMemcpyAsync(&d_peer->n_in, &c->h_peer->n_in, 16, cudaMemcpyHostToDevice, c->stream));        // n_in + vault_epoch + epoch
        }
        if (c->streaming && !c->stream_input_issued)
        {
            // the DMA front must not start moving ctl->in_ready before the control block of this cycle is in place
            QSB_CUDA(cudaEventRecord(c->ev_ctl, c->stream));
            QSB_CUDA(cudaStreamWaitEvent(c->stream_in, c->ev_ctl, 0));
        }
        QSB_CUDA(cudaEventRecord(c->ev0, c->stream));
        if (c->h_ctl->inflight > 0 || c->peer_on)        // peer mode: every rank takes part in every launch (arrivals, termination)
        {
            if (c->event_mode)
            {
                if (c->opt.validation) launch_track_event_validation(a, c->grid, c->stream);
                else                   launch_track_event_fast(a, c->grid, c->stream);
            }
            else if (c->opt.validation) launch_track_validation(a, c->grid, c->block, c->stream);
            else                        launch_track_fast(a, c->grid, c->block, c->stream);
            QSB_CUDA(cudaGetLastError());
            ++n_launch; c->launches++;
        }
        QSB_CUDA(cudaEventRecord(c->ev1, c->stream));
        if (c->streaming)
        {
            static const bool trace = std::getenv("QSB_TRACE") != nullptr;
            const auto t0 = std::chrono::steady_clock::now();
            issueStreamInput(c);                                    // H2D of the host vault runs under the kernel
            const auto t1 = std::chrono::steady_clock::now();
            while (cudaEventQuery(c->ev1) == cudaErrorNotReady) serviceCensus(c, false);   // D2H of finished census chunks too
            QSB_CUDA(cudaGetLastError());
            if (trace)
                std::fprintf(stderr, "[qsb] track(streamed): input enqueue %.2f ms, kernel wait %.2f ms, census records already on their way %llu\n",
                             std::chrono::duration<double, std::milli>(t1 - t0).count(),
                             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count(), c->census_copied);
        }
        pullControl(c);
        QSB_CUDA(cudaEventSynchronize(c->ev1));
        float ms = 0;
        QSB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (c->peer_on)
        {
            QSB_CUDA(cudaMemcpy(c->h_peer, c->peer_block, sizeof(PeerControl), cudaMemcpyDeviceToHost));
            c->h_ctl->inflight = c->h_peer->inflight;
            c->h_ctl->tail = c->h_peer->tail;
            if (c->h_peer->overflow == c->peer_epoch) c->h_ctl->overflow |= 1u;
            static const bool trace_peer = std::getenv("QSB_TRACE") != nullptr;
            if (trace_peer)
            {
                std::fprintf(stderr, "[qsb] rank %d peer launch %u: kernel %.3f ms, first idle after %.3f ms, global termination seen after %.3f ms, tail %llu; "
                             "send_advance: %llu calls, %.1f Mcycles over all warps; start-up wait %.3f ms\n",
                             c->my_rank, c->peer_epoch, ms, c->h_peer->first_idle_ns * 1e-6, c->h_peer->done_ns * 1e-6, c->h_peer->tail,
                             c->h_peer->send_calls, c->h_peer->send_cycles * 1e-6, c->h_peer->startup_wait_ns * 1e-6);
            }
            if (c->h_peer->abort == c->peer_epoch)
            {
                c->error = "peer exchange abandoned: a rank's watchdog expired before global termination (a rank that never launched, or lost particles)";
                return (int)QSB_ERR_INTERNAL;
            }
        }
        if (c->h_ctl->inflight != 0 && !c->h_ctl->overflow)
        { c->error = "tracking kernel ended with histories in flight"; return (int)QSB_ERR_INTERNAL; }
        {
            static const bool trace_track = std::getenv("QSB_TRACE") != nullptr;
            if (trace_track && !c->peer_on) std::fprintf(stderr, "[qsb] rank %d track: kernel %.3f ms, tail %llu\n", c->my_rank, ms, c->h_ctl->tail);
        }
        c->consumed = std::min<unsigned long long>(c->h_ctl->tail, c->n_in_aos + a.proc.capacity);
        if (stats)
        {
            stats->n_processed = c->consumed;
            stats->n_census = c->h_ctl->census_count;
            unsigned long long sent = 0;
            for (int r = 0; r < c->n_ranks; ++r) sent += c->h_ctl->send_count[r];
            stats->n_sent = sent;
            stats->n_launches = n_launch;
            stats->device_ms = ms;
        }
        if (c->h_ctl->overflow)
        {
            char msg[160];
            if (c->h_ctl->overflow & 8u)
            { c->error = "the tracking kernel's watchdog expired with histories still in flight (internal error)"; return (int)QSB_ERR_INTERNAL; }
            std::snprintf(msg, sizeof msg, "fixed-capacity storage overflowed (mask %u: 1 processing vault, 2 census vault, 4 send slab); "
                          "raise qsb_options.particle_capacity / send_capacity", c->h_ctl->overflow);
            c->error = msg;
            return (int)QSB_ERR_CAPACITY;
        }
        if (c->h_ctl->bad_group)
        { c->error = "a particle's energy lies above the last energy-group edge (eMax below the fission spectrum's 20 MeV?); the reference's behaviour there is undefined"; return (int)QSB_ERR_INTERNAL; }
        if (c->h_ctl->bad_reaction)
        { c->error = "a collision selected no reaction (cross-section table inconsistent)"; return (int)QSB_ERR_INTERNAL; }
        return (int)QSB_OK;
    });
}

int qsb_stream_end(qsb_ctx* c, uint64_t* n_census)
{
    if (!n_census) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        if (!c->streaming) { c->error = "qsb_stream_end without qsb_stream_begin"; return (int)QSB_ERR_STATE; }
        static const bool trace = std::getenv("QSB_TRACE") != nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        pullControl(c);
        *n_census = c->h_ctl->census_count;
        const unsigned long long before = c->census_copied;
        serviceCensus(c, true);
        QSB_CUDA(cudaStreamSynchronize(c->stream_out));
        QSB_CUDA(cudaStreamSynchronize(c->stream_in));
        if (trace)
        {
            float in_ms = 0.f, k_ms = 0.f;
            if (c->n_in_aos && cudaEventElapsedTime(&in_ms, c->ev0, c->ev_in_done) != cudaSuccess) { cudaGetLastError(); in_ms = -1.f; }
            if (cudaEventElapsedTime(&k_ms, c->ev0, c->ev1) != cudaSuccess) { cudaGetLastError(); k_ms = -1.f; }
            std::fprintf(stderr, "[qsb] stream_end: %llu of %llu census records were left to copy, %.2f ms; last input chunk landed %.2f ms after the "
                         "kernel's start (%.1f GB/s), kernel ended after %.2f ms\n",
                         (unsigned long long)*n_census - before, (unsigned long long)*n_census,
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), in_ms,
                         in_ms > 0 ? c->n_in_aos * sizeof(qsb_base_particle) / (in_ms * 1e6) : 0.0, k_ms);
        }
        return (int)QSB_OK;
    });
}

int qsb_get_census_range(qsb_ctx* c, uint64_t first, qsb_base_particle* out, uint64_t count)
{
    if (count && !out) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        if (!c->streaming) { c->error = "qsb_get_census_range: the census of this cycle is not in record (streamed) form"; return (int)QSB_ERR_STATE; }
        pullControl(c);
        if (first + count > c->h_ctl->census_count) { c->error = "census range out of bounds"; return (int)QSB_ERR_ARG; }
        if (count == 0) return (int)QSB_OK;
        if (isPinnedHost(out))
        {
            QSB_CUDA(cudaMemcpyAsync(out, c->d_census_aos + first, count * sizeof(qsb_base_particle), cudaMemcpyDeviceToHost, c->stream_out));
            QSB_CUDA(cudaStreamSynchronize(c->stream_out));
        }
        else QSB_CUDA(cudaMemcpy(out, c->d_census_aos + first, count * sizeof(qsb_base_particle), cudaMemcpyDeviceToHost));
        return (int)QSB_OK;
    });
}

int qsb_track_host(qsb_ctx* c, const qsb_base_particle* in, uint64_t n_in, qsb_base_particle* census_out, uint64_t census_cap,
                   uint64_t* n_census, qsb_track_stats* stats)
{
    int rc = qsb_stream_begin(c, in, n_in, census_out, census_cap);
    if (rc != QSB_OK) return rc;
    rc = qsb_track(c, stats);
    if (rc != QSB_OK) return rc;
    return qsb_stream_end(c, n_census);
}

int qsb_get_diagnostics(qsb_ctx* c, uint64_t out[8])
{
    if (!out) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        pullControl(c);
        out[0] = c->h_ctl->slow_geometry; out[1] = c->h_ctl->geometry_mismatch; out[2] = c->h_ctl->n_lookups;
        out[3] = (uint64_t)c->im.compact; out[4] = (uint64_t)c->regs; out[5] = (uint64_t)c->blocks_per_sm;
        out[6] = (uint64_t)c->grid; out[7] = c->h_ctl->tail;
        return (int)QSB_OK;
    });
}

int qsb_census_count(qsb_ctx* c, uint64_t* n)
{
    if (!n) return QSB_ERR_ARG;
    return guarded(c, [&]() { pullControl(c); *n = c->h_ctl->census_count; return (int)QSB_OK; });
}

int qsb_get_census(qsb_ctx* c, qsb_base_particle* aos, uint64_t cap, uint64_t* n_out)
{
    if (!n_out) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        pullControl(c);
        const uint64_t n = c->h_ctl->census_count;
        *n_out = n;
        if (n > cap || (n && !aos)) { c->error = "census buffer too small"; return (int)QSB_ERR_CAPACITY; }
        if (c->streaming)
        {
            // this cycle's census was written in record form (qsb_stream_begin): one plain copy
            if (n) QSB_CUDA(cudaMemcpy(aos, c->d_census_aos, n * sizeof(qsb_base_particle), cudaMemcpyDeviceToHost));
            return (int)QSB_OK;
        }
        const VaultView& v = c->vault[1 - c->proc];
        const size_t chunk = c->staging_records;
        for (uint64_t done = 0; done < n; done += chunk)
        {
            const uint64_t m = std::min<uint64_t>(chunk, n - done);
            const int grid = (int)std::min<uint64_t>((m + 255) / 256, 4096);
            soa_to_aos_kernel<<<grid, 256, 0, c->stream>>>(v, done, m, (qsb_base_particle*)c->staging, c->d_domain_offset, c->im.n_domains);
            c->launches++;
            QSB_CUDA(cudaMemcpyAsync(c->h_staging, c->staging, m * sizeof(qsb_base_particle), cudaMemcpyDeviceToHost, c->stream));
            QSB_CUDA(cudaStreamSynchronize(c->stream));
            std::memcpy(aos + done, c->h_staging, m * sizeof(qsb_base_particle));
        }
        QSB_CUDA(cudaGetLastError());
        return (int)QSB_OK;
    });
}

int qsb_get_balance(qsb_ctx* c, uint64_t out[QSB_BAL_COUNT])
{
    if (!out) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        pullControl(c);
        for (int i = 0; i < QSB_BAL_COUNT; ++i) out[i] = c->h_ctl->balance[i];
        return (int)QSB_OK;
    });
}

int qsb_get_scalar_flux(qsb_ctx* c, double* out)
{
    if (!out) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        QSB_CUDA(cudaMemcpyAsync(out, c->flux, (size_t)c->im.n_cells * c->im.n_groups * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        QSB_CUDA(cudaStreamSynchronize(c->stream));
        return (int)QSB_OK;
    });
}

int qsb_scalar_flux_sum(qsb_ctx* c, double* sum)
{
    if (!sum) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        const unsigned long long n = (unsigned long long)c->im.n_cells * c->im.n_groups;
        const int blocks = (int)std::min<unsigned long long>(1024, (n + 255) / 256);
        flux_partial_sum_kernel<<<blocks, 256, 0, c->stream>>>(c->flux, n, c->flux_partial);
        c->launches++;
        QSB_CUDA(cudaGetLastError());
        QSB_CUDA(cudaMemcpyAsync(c->h_partial, c->flux_partial, blocks * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        QSB_CUDA(cudaStreamSynchronize(c->stream));
        double s = 0.0;
        for (int i = 0; i < blocks; ++i) s += c->h_partial[i];
        *sum = s;
        return (int)QSB_OK;
    });
}

int qsb_send_counts(qsb_ctx* c, uint64_t* counts)
{
    if (!counts) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        pullControl(c);
        for (int r = 0; r < c->n_ranks; ++r) counts[r] = c->h_ctl->send_count[r];
        return (int)QSB_OK;
    });
}

int qsb_send_slab(qsb_ctx* c, int peer, void** device_ptr, uint64_t* n_records)
{
    if (!device_ptr || !n_records) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        if (peer < 0 || peer >= c->n_ranks) return (int)QSB_ERR_ARG;
        *device_ptr = c->sends + (size_t)peer * c->send_capacity;
        *n_records = c->h_ctl->send_count[peer];
        return (int)QSB_OK;
    });
}

int qsb_clear_sends(qsb_ctx* c)
{
    return guarded(c, [&]() {
        for (int r = 0; r < 64; ++r) c->h_ctl->send_count[r] = 0;
        QSB_CUDA(cudaMemsetAsync(c->d_ctl->send_count, 0, sizeof(c->h_ctl->send_count), c->stream));
        QSB_CUDA(cudaStreamSynchronize(c->stream));
        return (int)QSB_OK;
    });
}


int qsb_peer_export(qsb_ctx* c, void* handle, uint64_t* vault_capacity)
{
    if (!handle) return QSB_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == QSB_PEER_HANDLE_BYTES, "IPC handle size");
    static_assert(kMaxPeers == QSB_MAX_PEERS, "peer limit");
    static_assert(sizeof(PeerControl) <= kVaultHeaderBytes, "vault header");
    return guarded(c, [&]() {
        if (c->im.n_domains < 1 || c->im.n_domains > kMaxDomainsPerRank)
        { c->error = "peer exchange: at most " + std::to_string(kMaxDomainsPerRank) + " domains per rank"; return (int)QSB_ERR_STATE; }
        if (!c->peer_block)
        {
            c->peer_block = c->vault_base[0];
            QSB_CUDA(cudaMallocHost((void**)&c->h_peer, sizeof(PeerControl)));
            std::memset(c->h_peer, 0, sizeof(PeerControl));
            // where this rank's domains start in its flat cell index space: read by depositing peers (peer_destination_cell)
            c->h_peer->n_domains = c->im.n_domains;
            for (int d = 0; d < c->im.n_domains; ++d) c->h_peer->domain_offset[d] = c->host_domain_offset[d];
            PeerControl* d_peer = reinterpret_cast<PeerControl*>(c->peer_block);
            QSB_CUDA(cudaMemcpy(&d_peer->n_domains, &c->h_peer->n_domains, sizeof(int) * (1 + kMaxDomainsPerRank), cudaMemcpyHostToDevice));
        }
        cudaIpcMemHandle_t h;
        QSB_CUDA(cudaIpcGetMemHandle(&h, c->peer_block));
        std::memcpy(handle, &h, sizeof h);
        if (vault_capacity) *vault_capacity = c->vault[0].capacity;
        return (int)QSB_OK;
    });
}

int qsb_peer_connect(qsb_ctx* c, const void* handles, int n_ranks, double watchdog_seconds)
{
    if (!handles) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        if (!c->peer_block) { c->error = "qsb_peer_connect before qsb_peer_export"; return (int)QSB_ERR_STATE; }
        if (n_ranks != c->n_ranks || n_ranks > kMaxPeers)
        { c->error = "qsb_peer_connect: rank count must equal the image's and be at most QSB_MAX_PEERS"; return (int)QSB_ERR_ARG; }
        if (c->peer_on) return (int)QSB_OK;
        for (int r = 0; r < n_ranks; ++r)
        {
            if (r == c->my_rank) { c->peer_base[r] = c->peer_block; continue; }
            cudaIpcMemHandle_t h;
            std::memcpy(&h, (const char*)handles + (size_t)r * QSB_PEER_HANDLE_BYTES, sizeof h);
            void* p = nullptr;
            const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
            {
                cudaGetLastError();
                for (int q = 0; q < r; ++q) if (q != c->my_rank && c->peer_base[q]) { cudaIpcCloseMemHandle(c->peer_base[q]); c->peer_base[q] = nullptr; }
                c->error = std::string("cudaIpcOpenMemHandle failed for rank ") + std::to_string(r) + ": " + cudaGetErrorString(e);
                return (int)QSB_ERR_CUDA;
            }
            c->peer_base[r] = (char*)p;
        }
        // does any rank own several domains?  (every rank wrote its header in qsb_peer_export, before its handle left)
        c->peer_multi_domain = c->im.n_domains > 1;
        for (int r = 0; r < n_ranks; ++r)
        {
            if (r == c->my_rank) continue;
            int nd = 0;
            QSB_CUDA(cudaMemcpy(&nd, &reinterpret_cast<PeerControl*>(c->peer_base[r])->n_domains, sizeof(int), cudaMemcpyDeviceToHost));
            if (nd < 1 || nd > kMaxDomainsPerRank) { c->error = "qsb_peer_connect: rank " + std::to_string(r) + " has not exported its domain table"; return (int)QSB_ERR_STATE; }
            if (nd > 1) c->peer_multi_domain = true;
        }
        c->watchdog_ns = (unsigned long long)((watchdog_seconds > 0 ? watchdog_seconds : 60.0) * 1e9);
        c->peer_on = std::getenv("QSB_DEBUG_PEER_MAP_ONLY") == nullptr;      // experiment: mappings in place, exchange left to the caller
        return (int)QSB_OK;
    });
}

int qsb_peer_diagnostics(qsb_ctx* c, uint64_t out[8])
{
    if (!out) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        if (!c->h_peer) { c->error = "qsb_peer_diagnostics before qsb_peer_export"; return (int)QSB_ERR_STATE; }
        unsigned long long sent = 0;
        for (int r = 0; r < c->n_ranks; ++r) sent += c->h_ctl->send_count[r];
        out[0] = c->h_peer->first_idle_ns; out[1] = c->h_peer->done_ns; out[2] = c->h_peer->send_cycles; out[3] = c->h_peer->send_calls;
        out[4] = c->h_peer->startup_wait_ns; out[5] = c->h_peer->tail; out[6] = sent; out[7] = c->h_peer->bulk_done_ns;
        return (int)QSB_OK;
    });
}

int qsb_peer_disconnect(qsb_ctx* c)
{
    if (!c) return QSB_ERR_ARG;
    cudaSetDevice(c->device);
    for (int r = 0; r < kMaxPeers; ++r)
    {
        if (r != c->my_rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
        c->peer_base[r] = nullptr;
    }
    c->peer_on = false;
    return QSB_OK;
}


int qsb_fluence_accumulate(qsb_ctx* c)
{
    return guarded(c, [&]() {
        if (!c->in_cycle) { c->error = "qsb_fluence_accumulate before the first cycle"; return (int)QSB_ERR_STATE; }
        if (!c->fluence)
        {
            c->fluence = devAlloc<double>((size_t)c->im.n_cells, c->owned);
            QSB_CUDA(cudaMemsetAsync(c->fluence, 0, (size_t)c->im.n_cells * sizeof(double), c->stream));
        }
        const int grid = std::min((c->im.n_cells + 7) / 8, c->sm_count * 8);
        fluence_accumulate_kernel<<<grid, 256, 0, c->stream>>>(c->flux, c->im.n_cells, c->im.n_groups, c->fluence);
        c->launches++;
        QSB_CUDA(cudaGetLastError());
        return (int)QSB_OK;
    });
}

int qsb_census_energy_spectrum(qsb_ctx* c, uint64_t* counts, uint64_t n_counts)
{
    if (!counts) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        const int n_edges = c->im.n_groups + 1;
        if (n_counts != (uint64_t)n_edges) { c->error = "qsb_census_energy_spectrum: n_counts must be n_groups + 1"; return (int)QSB_ERR_ARG; }
        if ((size_t)n_edges * sizeof(unsigned int) > 48 * 1024) { c->error = "qsb_census_energy_spectrum: more than 12287 energy groups"; return (int)QSB_ERR_ARG; }
        if (!c->in_cycle) { std::memset(counts, 0, n_counts * sizeof(uint64_t)); return (int)QSB_OK; }
        pullControl(c);
        const unsigned long long n = std::min<unsigned long long>(c->h_ctl->census_count, c->vault[1 - c->proc].capacity);
        if (!c->d_spectrum) c->d_spectrum = devAlloc<unsigned long long>((size_t)n_edges, c->owned);
        QSB_CUDA(cudaMemsetAsync(c->d_spectrum, 0, (size_t)n_edges * sizeof(unsigned long long), c->stream));
        if (n)
        {
            const double* energy = c->streaming ? reinterpret_cast<const double*>(c->d_census_aos) + 6 : c->vault[1 - c->proc].energy;
            const int stride = c->streaming ? (int)(sizeof(qsb_base_particle) / sizeof(double)) : 1;
            const int grid = (int)std::min<unsigned long long>((n + 255) / 256, (unsigned long long)c->sm_count * 8);
            census_energy_histogram_kernel<<<grid, 256, (size_t)n_edges * sizeof(unsigned int), c->stream>>>(energy, n, stride, c->im.energies, n_edges, c->d_spectrum);
            c->launches++;
            QSB_CUDA(cudaGetLastError());
        }
        static_assert(sizeof(uint64_t) == sizeof(unsigned long long), "counter width");
        QSB_CUDA(cudaMemcpyAsync(counts, c->d_spectrum, (size_t)n_edges * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        QSB_CUDA(cudaStreamSynchronize(c->stream));
        return (int)QSB_OK;
    });
}

int qsb_get_fluence(qsb_ctx* c, double* out)
{
    if (!out) return QSB_ERR_ARG;
    return guarded(c, [&]() {
        if (!c->fluence) { std::memset(out, 0, (size_t)c->im.n_cells * sizeof(double)); return (int)QSB_OK; }
        QSB_CUDA(cudaMemcpyAsync(out, c->fluence, (size_t)c->im.n_cells * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        QSB_CUDA(cudaStreamSynchronize(c->stream));
        return (int)QSB_OK;
    });
}

} // extern "C"
