// device_types.cuh -- plain structs shared by the device context (device_ctx.cu) and the tracking
// kernels (track_kernels.cu).  Everything here is a view onto memory owned by qsb_ctx.
#ifndef QSB_DEVICE_TYPES_CUH
#define QSB_DEVICE_TYPES_CUH

#include <cstdint>
#include "../../../include/qsb.h"

namespace qsb {

// Compact per-cell record, 64 bytes (two sectors), everything the common tracking path reads about a cell.
// The mesh is always a uniform brick grid (reference: GlobalFccGrid), so a cell's 24 facet planes are
// axis aligned: the one non-zero normal component is +-1 or +-(1 - 2^-53) and D is the face coordinate
// displaced by -1/0/+1 ulp (rounding of src/MC_Facet_Geometry.hh:18-39).  code[f] records exactly that:
// bit0 = |normal| is 1 - 2^-53, bits1-2 = ulp displacement of |D| (0: none, 1: +1, 2: -1); the node
// coordinates themselves are index * cell size, so the full-precision plane is rebuilt in registers.
// The host verifies bit-for-bit that every facet decodes to the reference plane before enabling it.
struct __align__(16) CellRec
{
    uint16_t ix, iy, iz;        // global cell indices
    uint8_t  material;
    uint8_t  flags;
    uint32_t events;            // 4 bits per face: QSB_ADJ_*
    uint32_t pad;
    int32_t  adj[6];            // per face: adjacent flat cell (on-processor) / neighbour-domain cell (off-processor)
    uint8_t  code[24];          // per facet, see above
};
static_assert(sizeof(CellRec) == 64, "CellRec layout");

// read-only problem image in HBM (see include/qsb.h: qsb_image for the meaning of each array)
struct DevImage
{
    int n_cells, n_groups, n_materials, max_react, n_domains, n_ranks;
    int compact;                    // 1: CellRec codes reproduce every plane bit-for-bit -> fast geometry path enabled
    double dx, dy, dz;              // cell size (global_l / global_n, the host grid's own expression)
    double margin;                  // robustness margin of the fast geometry path: 1e-6 * smallest cell edge
    double inv_hx, inv_hy, inv_hz;  // 2/dx, 2/dy, 2/dz (filter arithmetic only)
    float group_log2_lo, group_inv_dlog2; // energy-group index estimate: (log2(E) - log2_lo) * inv_dlog2 (estimate only)
    const CellRec* cells;           // [n_cells]
    const double2* xs_pair;         // [n_materials*n_groups] {total, 1/total}
    const double4* planes;          // [n_cells*24] {A,B,C,D}, 32-byte aligned records
    const double*  nodes;           // [n_cells*42]
    const int*     face_adj_cell;   // [n_cells*6]
    const uint8_t* face_event;      // [n_cells*6]
    const int*     face_adj_domain; // [n_cells*6]
    const int*     face_nbr_rank;   // [n_cells*6]
    const int*     cell_material;   // [n_cells]
    const int*     domain_cell_offset; // [n_domains+1]
    // "hot block": one contiguous allocation covered by the persisting-L2 access-policy window
    const double*  energies;        // [n_groups+1]
    const double*  xs_total;        // [n_materials*n_groups]
    const double*  xs_react;        // [n_materials*n_groups*max_react]
    const double*  mat_inv_mass;    // [n_materials] 1.0 / mass (the reference forms this quotient per scatter, src/NuclearData.cc:66)
    const double*  mat_nu_bar;      // [n_materials]
    const int*     mat_n_iso;       // [n_materials]
    const int*     mat_n_react;     // [n_materials]
    const uint8_t* mat_react_type;  // [n_materials*max_react]
    const uint8_t* mat_periodic;    // [n_materials]
    // ---- derived, optional fast paths (set up and VERIFIED against the arrays above by qsb_create) ----
    // brick: every on-processor transit leads to cell + {+sx, -sx, +sy, -sy, +sz, -sz}[face] (one lexicographically numbered
    // brick per rank): the neighbour is computed instead of read from the cell record, and what the tracking loop needs of
    // the new cell -- face events and material, one word -- comes from a 4-byte-per-cell array (1 MB at 64^3 cells instead
    // of a 16 MB record array); its grid indices follow from the old ones.
    int brick;                      // 1: the computed-neighbour path is valid for this image
    int brick_stride[3];            // sx, sy, sz
    const uint32_t* cell_info;      // [n_cells] events (6 x 4 bits) | material << 24
    // several ranks: 1 for cells within a few cells of a face that leads to another rank (breadth-first over the on-rank
    // adjacency from the cells that have such a face); nullptr on a single rank.  See TrackArgs::prio_list.
    const uint8_t* cell_near;
    // compact reaction table: when every material is periodic (one reaction table shared by its isotopes) only the first
    // isotope's rows are ever read: [n_materials][n_groups][compact_react] instead of rows of max_react doubles
    const double*  xs_compact;      // nullptr: not available
    int compact_react;
};

// SoA particle vault: one array per MC_Base_Particle field (src/MC_Base_Particle.hh:75-92), cell is the
// FLAT cell index, dir* carry the direction cosine of arrivals from other ranks (NaN = derive from
// velocity, the reference's MC_Particle(const MC_Base_Particle&) behaviour).
struct VaultView
{
    double *x, *y, *z, *vx, *vy, *vz, *energy, *weight, *ttc, *age, *nmfp, *nseg;
    double *dirx, *diry, *dirz;
    unsigned long long *seed, *id;
    unsigned long long *check;      // peer deposits only: XOR of the record's words and a per-launch salt (the record validates itself)
    int *cell;
    int4 *tags;                     // {last_event, num_collisions, breed, species}
    uint32_t *ready;                // == epoch once the slot is fully written (processing vault only); epoch | kArrivalBit: a
                                    // peer's deposit is on its way into the slot -- complete when `check` matches the record
    unsigned long long capacity;
};

// exchange record for boundary-crossing particles: MC_Base_Particle + direction cosine (160 bytes)
struct ExchangeRecord
{
    qsb_base_particle p;
    double dir[3];
};

// queue control + tallies, one struct in device memory so a single small copy brings the state back
struct DevControl
{
    // The four queue counters are hit by atomics from every warp on the GPU (together ~20 M per CORAL2 cycle); each sits in
    // its own 256-byte block so that they are served by different L2 slices instead of queueing on one atomic unit.
    unsigned long long head;            // next unclaimed ticket of the processing queue
    unsigned long long pad0[31];
    unsigned long long tail;            // tickets allocated (streamed records + vault slots: initial, arrivals, secondaries)
    unsigned long long pad1[31];
    unsigned long long census_count;
    unsigned long long pad2[31];
    unsigned long long inflight;        // histories created and not yet finished (queued + running); 0 = cycle drained
    unsigned long long pad3[31];
    unsigned long long in_ready;        // host-buffer streaming: input records [0, in_ready) have landed in HBM
    unsigned long long pad4[31];
    unsigned long long head_in;         // event kernel, host-buffer streaming: next unclaimed ticket of the INPUT queue [0, n_in); `head` then
    unsigned long long pad5[31];        // serves the vault slots only (tickets >= n_in), so secondaries do not queue behind the whole input
    unsigned long long arr_head;        // event kernel, peer mode: next unclaimed slot of the arrival region (its tail: PeerControl::arr_tail)
    unsigned long long pad6[31];
    unsigned long long prio_head;       // event kernel, peer mode: next unclaimed entry of the boundary-first list (TrackArgs::prio_list)
    unsigned long long pad7[31];
    unsigned long long prio_count;      // entries in that list (written by boundary_list_kernel before the tracking launch)
    unsigned long long pad8[31];
    unsigned long long slow_geometry;   // segments that took the full 24-facet path
    unsigned long long geometry_mismatch; // check mode: fast and full path disagreed (must stay 0)
    unsigned long long balance[QSB_BAL_COUNT];
    unsigned long long n_lookups;       // diagnostics
    unsigned int overflow;              // bit0 processing vault, bit1 census vault, bit2 send slab
    unsigned int bad_reaction;          // collisions where no reaction was selected (reference: unreachable)
    unsigned int epoch;
    unsigned int bad_group;             // particles whose energy lies above the last group edge (the reference indexes out of bounds there)
    unsigned long long send_count[64];  // per peer rank
};

// Peer exchange over NVLink (one process per GPU, all GPUs of one NVSwitch domain).  Every rank exports ONE allocation
// through CUDA IPC: this control block followed by its processing vault (all SoA arrays in one allocation, see
// vault_view).  A tracking kernel whose particle crosses onto a peer's domain takes a slot of the PEER's processing vault
// with a system-scope atomic on the peer's tail counter, raises the peer's in-flight count, stores the particle into the
// peer's SoA arrays through NVLink and releases the slot's ready word -- to the peer's tracking warps the arrival looks
// like a fission secondary that appeared in their ticket queue; nothing on the receiving side copies or converts.
// `sent` / `received` are monotonic over the life of the context; `tail`, `inflight`, `n_in`, `vault_epoch` belong to the
// launch named by `epoch` (the host writes them, then `epoch`, in stream order before the launch; a sender waits for the
// peer's epoch before it touches them).  Global termination is decided on the devices: peer_service_loop in track_kernels.cu.
constexpr int kMaxPeers = 8;
// Event kernel, peer mode: particles deposited by other GPUs do not queue behind this GPU's own population and secondaries
// (one FIFO made every hop of a particle that crosses back and forth wait for the bulk of the cycle, and the chains of hops
// then ran one after the other on nearly empty GPUs at the end).  They get a region of their own at the top of the
// processing vault -- the last capacity / 8 slots -- with its own tail (PeerControl::arr_tail, raised by the senders) and
// head (DevControl::arr_head), and LOAD serves it first.  An arrival's ticket is its slot number in the region + kArrivalTicket.
#ifndef QSB_OPT_ARRIVAL_QUEUE
#define QSB_OPT_ARRIVAL_QUEUE 1
#endif
constexpr unsigned long long kArrivalTicket = 1ull << 62;
// Event kernel, peer mode: BOUNDARY FIRST.  A particle that crosses to another GPU starts a chain of hops (it may cross back
// and forth: 11 exchange rounds in the NCCL mode of Coral2_P1), each hop a whole history segment chain on the other GPU; a
// chain started at the end of the cycle is tracked on nearly empty GPUs, one hop after the other, and every rank waits for
// it (measured on 2 GPUs: the kernel ran 2.3 ms beyond the 12.1 ms of its own work).  So the histories that can reach
// another rank -- those starting within a few cells of an off-rank face, DevImage::cell_near -- are tracked FIRST: a small
// kernel lists their vault slots (boundary_list_kernel), LOAD serves that list before the vault queue, and the vault queue
// skips the slots the list holds (same predicate, so no slot is tracked twice and none is lost).  A list ticket is its index
// + kPrioTicket.
#ifndef QSB_OPT_BOUNDARY_FIRST
#define QSB_OPT_BOUNDARY_FIRST 1
#endif
constexpr unsigned long long kPrioTicket = 1ull << 61;
constexpr int kMaxDomainsPerRank = 64;
struct PeerControl
{
    unsigned long long sent;            // deposits started TOWARDS this GPU (remote atomics, while the sender still counts the history)
    unsigned long long pad0[31];
    unsigned long long received;        // deposits counted into this GPU's in-flight count (remote atomics, after `inflight`)
    unsigned long long pad1[31];
    unsigned long long inflight;        // histories queued or running on this GPU (the kernel's in-flight counter in peer mode)
    unsigned long long pad2[31];
    unsigned long long tail;            // tickets allocated (the kernel's queue tail in peer mode; history kernel: remote senders add to it)
    unsigned long long pad3[31];
    unsigned long long arr_tail;        // event kernel: slots of the ARRIVAL region handed out to depositing peers this cycle (remote atomics)
    unsigned long long pad3b[31];
    unsigned long long n_in;            // this launch: tickets below n_in are streamed host records, SoA slot = ticket - n_in
    unsigned int vault_epoch;           // this launch: value of a slot's ready word once it is fully written
    unsigned int epoch;                 // number of the peer-mode launch the words above belong to
    unsigned int done;                  // == epoch: this GPU has seen global termination of that launch
    unsigned int abort;                 // == epoch: some GPU gave up (watchdog); everybody leaves
    unsigned int overflow;              // == epoch: a sender found this GPU's processing vault full
    unsigned int pad4;
    unsigned long long first_idle_ns;   // diagnostics of the last launch: kernel start -> this GPU first had nothing queued or running,
    unsigned long long done_ns;         //                                  kernel start -> global termination seen
    unsigned long long send_cycles;     //   SM cycles warps spent inside send_advance, summed over warps; calls; start-up wait cycles (block 0)
    unsigned long long send_calls;
    unsigned long long startup_wait_ns;
    unsigned long long bulk_done_ns;    //   kernel start -> a warp first found this GPU's own vault queue empty (what follows is arrivals and their chains)
    unsigned int pad5[12];
    // written once by qsb_peer_export: where each of this rank's domains starts in its flat cell index space (a depositing
    // peer knows the destination as (rank-local domain, cell) and stores the flat cell, src/initMC.cc:256-259: a rank may own
    // several domains)
    int n_domains;
    int domain_offset[kMaxDomainsPerRank];
    int pad6[3];
};
static_assert(sizeof(PeerControl) == 1152 + 256 + 16 + 4 * kMaxDomainsPerRank, "PeerControl layout");
constexpr size_t kVaultHeaderBytes = 2048;     // PeerControl sits at the head of the processing vault's allocation

// the SoA arrays of a vault inside one allocation: 18 eight-byte arrays, tags, cell, ready (capacity is a multiple of 32)
__host__ __device__ inline VaultView vault_view(char* base, unsigned long long cap)
{
    VaultView v;
    double* d = reinterpret_cast<double*>(base + kVaultHeaderBytes);
    v.x = d; v.y = d + cap; v.z = d + 2 * cap; v.vx = d + 3 * cap; v.vy = d + 4 * cap; v.vz = d + 5 * cap;
    v.energy = d + 6 * cap; v.weight = d + 7 * cap; v.ttc = d + 8 * cap; v.age = d + 9 * cap; v.nmfp = d + 10 * cap; v.nseg = d + 11 * cap;
    v.dirx = d + 12 * cap; v.diry = d + 13 * cap; v.dirz = d + 14 * cap;
    v.seed = reinterpret_cast<unsigned long long*>(d + 15 * cap);
    v.id = reinterpret_cast<unsigned long long*>(d + 16 * cap);
    v.check = reinterpret_cast<unsigned long long*>(d + 17 * cap);
    v.tags = reinterpret_cast<int4*>(d + 18 * cap);
    v.cell = reinterpret_cast<int*>(d + 20 * cap);
    v.ready = reinterpret_cast<uint32_t*>(v.cell + cap);
    v.capacity = cap;
    return v;
}
inline size_t vault_bytes(unsigned long long cap) { return kVaultHeaderBytes + (size_t)cap * (20 * 8 + 4 + 4); }
constexpr uint32_t kArrivalBit = 0x80000000u;
__host__ __device__ inline unsigned long long deposit_salt(uint32_t vault_epoch) { return ((unsigned long long)vault_epoch + 1ull) * 0x9E3779B97F4A7C15ull; }

struct TrackArgs
{
    DevImage im;
    VaultView proc;
    VaultView census;
    ExchangeRecord* sends;              // [n_ranks][send_capacity]
    unsigned long long send_capacity;
    DevControl* ctl;
    double* flux;                       // [n_cells][n_groups]
    double dt;
    unsigned long long ready_prefix;    // SoA slots below this index were written by the host side
    // host-buffer streaming (qsb_track_host): the first n_in tickets are AoS records of the host vault, DMA-copied chunk by
    // chunk into in_aos while the kernel runs (ctl->in_ready says how far); SoA slot = ticket - n_in for the rest.  Census
    // records are written as AoS into census_aos; every full chunk of 2^census_chunk_shift records is announced to the host
    // through a flag in mapped pinned memory so that its D2H copy overlaps the tracking still going on.
    const qsb_base_particle* in_aos;
    unsigned long long n_in;
    qsb_base_particle* census_aos;      // nullptr: census goes to the SoA vault
    unsigned int* census_chunk_done;    // [chunks] records completed per chunk (device memory)
    unsigned int* host_chunk_flags;     // [chunks] mapped pinned host memory: == epoch once the chunk is complete
    unsigned int census_chunk_shift;
    uint32_t epoch;                     // value of a slot's ready word once it is fully written this cycle
    int check_mode;                     // bit 0: evaluate both geometry paths, bit 1: both reaction selections; count disagreements
    unsigned long long* inflight;       // the in-flight counter and the queue tail: ctl's, or the exported PeerControl's in peer mode
    unsigned long long* tail;
    // peer exchange (peer_mode != 0): base of every rank's exported allocation (own rank: the local pointer); every rank's
    // processing vault has this rank's capacity
    int peer_mode, my_rank;
    int peer_multi_domain;              // some rank owns more than one domain: a deposit adds the destination domain's offset (read from the peer)
    unsigned long long arrival_first, arrival_cap;  // the arrival region [arrival_first, arrival_first + arrival_cap) of every rank's vault (0: none)
    const uint32_t* prio_list;          // boundary-first list: vault slots (< prio_slots) whose cell is near another rank; nullptr: none
    unsigned long long prio_slots;      // the list was built over vault slots [0, prio_slots): the vault queue skips listed slots below it
    uint32_t peer_epoch;
    char* peer_base[kMaxPeers];
    unsigned long long watchdog_ns;     // give up (abort everywhere) when a launch has not terminated after this long
};

// cycleInit on the device (cycle_init_kernels.cu): last cycle's census + this cycle's source particles -> population control
// -> low-weight roulette -> processing vault, one kernel
struct CycleInitCounters
{
    unsigned long long n_out;           // records appended to the processing vault
    unsigned long long n_rr;            // particles killed by population control or the low-weight roulette (Balance::_rr)
    unsigned long long n_split;         // split copies made (Balance::_split)
    unsigned int overflow, pad;
};
struct CycleInitArgs
{
    VaultView src;                      // census vault of the previous cycle, records [0, n_carried)
    VaultView dst;                      // processing vault of this cycle
    unsigned long long n_carried, n_source;
    const int* source_offsets;          // [n_cells+1] prefix sum of the per-cell source counts
    const unsigned long long* source_tally; // [n_cells] the cells' running source counts before this cycle
    const unsigned long long* cell_id;  // [n_cells]
    const double* cell_volume;          // [n_cells]
    const double* nodes;                // [n_cells*42]
    int n_cells;
    double source_weight, e_min, e_max, dt;
    double factor;                      // population-control factor (1.0: none)
    double cutoff, weight_cutoff;       // low-weight roulette: relative cut-off (<= 0: off) and cut-off * source weight
    uint32_t epoch;
    CycleInitCounters* out;
};
void launch_cycle_init(const CycleInitArgs& a, int sm_count, cudaStream_t s);
void launch_source_tally_advance(unsigned long long* tally, const int* offsets, int n_cells, cudaStream_t s);

// launchers implemented twice in track_kernels.cu (validation: --fmad=false + strict math; fast)
void launch_track_validation(const TrackArgs& a, int grid, int block, cudaStream_t s);
void launch_track_fast(const TrackArgs& a, int grid, int block, cudaStream_t s);
void track_kernel_attributes_validation(int* regs, int* max_blocks_per_sm, int block);
void track_kernel_attributes_fast(int* regs, int* max_blocks_per_sm, int block);
// the event-based kernels (track_event_kernels.cu; tracking_mode bit 0): block shape and shared memory are compile-time
void launch_track_event_validation(const TrackArgs& a, int grid, cudaStream_t s);
void launch_track_event_fast(const TrackArgs& a, int grid, cudaStream_t s);
void track_event_kernel_attributes_validation(int* regs, int* max_blocks_per_sm, int* threads, int* smem_bytes, int* slots);
void track_event_kernel_attributes_fast(int* regs, int* max_blocks_per_sm, int* threads, int* smem_bytes, int* slots);

} // namespace qsb
#endif
