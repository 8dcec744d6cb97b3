// track_event_kernels.cu -- event-based cycle tracking (the default; tracking_mode bit 0 = 1 selects the history kernel).
//
// The history-based kernel (track_kernels.cu) keeps one history per lane in registers and regroups lanes inside a warp by
// what they need next; ncu puts it at 18.6 of 32 lanes per instruction with 160 registers of live particle state + event
// code.  Here the particles live in SHARED MEMORY -- structure of arrays over slots, ~175 bytes per particle, kWq slots
// owned by every WARP -- and a warp runs batches of ONE event type over its own slots:
//
//   LOAD       take a ticket of the processing-vault queue, redeem it when the slot is written, particle -> shared memory
//   SEGMENT    MC_Segment_Outcome + (for a facet crossing) MC_Facet_Crossing_Event         src/MC_Segment_Outcome.cc, ...
//   COLLISION  CollisionEvent (+ the outgoing trajectory of freshly loaded fission secondaries)   src/CollisionEvent.cc
//   CENSUS     store the record in the census vault                                               src/CycleTracking.cc:93-103
//   SEND       a particle that left the rank's domain: NVLink deposit on the peer / send slab     src/MC_Facet_Crossing_Event.cc:49-67
//
// A slot's state byte says which event it waits for.  Per iteration the warp picks one event type (one with >= 32 waiting
// slots if there is one: collisions first, they are the longest), gathers up to 32 such slots into a list with ballots,
// loads exactly the fields that event reads, runs the SAME device functions as the history kernel (track_physics.cuh: the
// arithmetic, and therefore every bit of every result, is shared), stores the fields the event may have changed and sets
// the slots' next states.  With 96 slots per warp every SEGMENT / COLLISION batch is full (31.2 of 32 lanes enter a segment
// batch, ncu); no particle state is live in registers across events; nothing is shared between warps -- no block barrier, no
// shared-memory atomics (a first version with block-wide queues and one __syncthreads per phase lost 19 % of its cycles at
// the barrier and ran 0.85x the history kernel; this one runs 1.17x, profiles/r02_event_kernel_shapes.txt).
//
// Everything outside the warp is the history kernel's: the ticket queue over the processing vault (fission secondaries
// are appended to the vault and picked up by whichever warp redeems their ticket), the global in-flight counter, the
// census append, the peer deposits and the termination protocol.  Results are independent of the scheduling: a
// history's random numbers come from its own stream and tallies are sums (tests: bit-identical census and balance).
#include <cstdlib>

#include "track_physics.cuh"

namespace qsb {
namespace {



// a block's particle slots: structure of arrays over N slots
template <int N>
struct SlotStore
{
    double x[N], y[N], z[N], alpha[N], beta[N], gamma[N];
    double energy[N], weight[N], ttc[N], age[N], nmfp[N], nseg[N], speed[N];
#if QSB_VALIDATION
    double vx[N], vy[N], vz[N];
#endif
    unsigned long long seed[N];
    unsigned long long id[N];              // LOAD state: the slot's ticket (kNoTicket: none yet)
    uint2 head01[N];                       // CellRec words 0-1: ix | iy << 16, iz | material << 16
    unsigned events[N];                    // CellRec word 2: 4 bits per face
    int cell[N];
    int num_collisions[N], breed[N], species[N];
    unsigned short group[N];
    unsigned char facet[N], last_event[N], flags[N];
};


// the balance counters a warp keeps (Counters of the history kernel), see the kernel's epilogue
enum { kTalSegments = 0, kTalCollisions, kTalAbsorbs, kTalFissions, kTalProduced, kTalEscapes, kTalCensus, kTalScatters, kTalLookups, kTalSlow, kTalMismatch };

template <class Store>
__device__ __forceinline__ void load_head(const Store& s, unsigned slot, Particle& p)
{
    const uint2 h = s.head01[slot];
    p.head = make_uint4(h.x, h.y, s.events[slot], 0u);
}
template <class Store>
__device__ __forceinline__ void store_head(Store& s, unsigned slot, const Particle& p)
{
    s.head01[slot] = make_uint2(p.head.x, p.head.y);
    s.events[slot] = p.head.z;
}

// every field of a particle, registers <-> slot (LOAD writes, CENSUS / SEND read)
template <class Store>
__device__ __forceinline__ void store_all(Store& s, unsigned slot, const Particle& p)
{
    s.x[slot] = p.x; s.y[slot] = p.y; s.z[slot] = p.z;
    s.alpha[slot] = p.alpha; s.beta[slot] = p.beta; s.gamma[slot] = p.gamma;
    s.energy[slot] = p.energy; s.weight[slot] = p.weight; s.ttc[slot] = p.ttc; s.age[slot] = p.age;
    s.nmfp[slot] = p.nmfp; s.nseg[slot] = p.nseg; s.speed[slot] = p.speed;
#if QSB_VALIDATION
    s.vx[slot] = p.vx; s.vy[slot] = p.vy; s.vz[slot] = p.vz;
#endif
    s.seed[slot] = p.seed; s.id[slot] = p.id;
    store_head(s, slot, p);
    s.cell[slot] = p.cell;
    s.num_collisions[slot] = p.num_collisions; s.breed[slot] = p.breed; s.species[slot] = p.species;
    s.group[slot] = (unsigned short)p.group;
    s.facet[slot] = (unsigned char)p.facet; s.last_event[slot] = (unsigned char)p.last_event;
}
template <class Store>
__device__ __forceinline__ void load_all(const Store& s, unsigned slot, Particle& p)
{
    p.x = s.x[slot]; p.y = s.y[slot]; p.z = s.z[slot];
    p.alpha = s.alpha[slot]; p.beta = s.beta[slot]; p.gamma = s.gamma[slot];
    p.energy = s.energy[slot]; p.weight = s.weight[slot]; p.ttc = s.ttc[slot]; p.age = s.age[slot];
    p.nmfp = s.nmfp[slot]; p.nseg = s.nseg[slot]; p.speed = s.speed[slot];
#if QSB_VALIDATION
    p.vx = s.vx[slot]; p.vy = s.vy[slot]; p.vz = s.vz[slot];
#else
    p.vx = p.vy = p.vz = 0.0;
#endif
    p.seed = s.seed[slot]; p.id = s.id[slot];
    load_head(s, slot, p);
    p.cell = s.cell[slot];
    p.num_collisions = s.num_collisions[slot]; p.breed = s.breed[slot]; p.species = s.species[slot];
    p.group = s.group[slot];
    p.facet = s.facet[slot]; p.last_event = s.last_event[slot];
    p.total_xs = 0.0;
}

// shape (csrc/Makefile: EVT_DEFS): particle slots per warp, warps per block, resident blocks per SM
#ifndef QSB_WQ_SLOTS_PER_WARP
#define QSB_WQ_SLOTS_PER_WARP 96
#endif
#ifndef QSB_WQ_WARPS
#define QSB_WQ_WARPS 4
#endif
#ifndef QSB_WQ_MIN_BLOCKS
#define QSB_WQ_MIN_BLOCKS 3
#endif
constexpr int kWq = QSB_WQ_SLOTS_PER_WARP;          // slots per warp
constexpr int kWqK = (kWq + 31) / 32;               // ... per lane of bookkeeping
constexpr int kWqWarps = QSB_WQ_WARPS;
constexpr int kWqThreads = 32 * kWqWarps;
constexpr int kWqSlots = kWq * kWqWarps;
// the service events as functions of their own (1, default) or inlined into the kernel (0)
#ifndef QSB_OPT_SERVICE_CALLS
#define QSB_OPT_SERVICE_CALLS 1
#endif
#if QSB_OPT_SERVICE_CALLS
#define QSB_WQ_SERVICE_FN __noinline__
#else
#define QSB_WQ_SERVICE_FN __forceinline__
#endif
#ifndef QSB_OPT_SERVICE_WATCHDOG
#define QSB_OPT_SERVICE_WATCHDOG 1
#endif
#ifndef QSB_OPT_SERVICE_ALL
#define QSB_OPT_SERVICE_ALL 0           // 1: a service phase always does everything (send, census, refill); 0: sends + the larger of census / refill
#endif
#ifndef QSB_WQ_SERVICE
#define QSB_WQ_SERVICE (QSB_WQ_SLOTS_PER_WARP / 2)
#endif
constexpr int kService = QSB_WQ_SERVICE;            // parked slots (census / send / empty) a warp lets gather before it services them:
                                                    // half of its slots (measured, Coral2_P1: 24 -> 15.6 ms, 32 -> 15.0, 48 -> 14.4)
static_assert(kWq % 4 == 0 && kWq >= 32 && kWq <= 256, "a warp's slots: at least a batch, slot numbers fit a byte");

enum { kStLoad = 0, kStSegment, kStCollision, kStTail, kStCensus, kStSend };

// A warp's scheduler state lives in shared memory, not in registers: the event code in between is where the register
// budget goes, and anything that stays live across it is spilled to local memory (measured: the spilled per-thread balance
// counters and slot counts alone were 15 % of all stall samples).  Everybody reads it at the top of an iteration (one
// broadcast load each), lane 0 updates it at the end of a batch.
enum { kNSeg = 0, kNCol, kNCen, kNSnd, kNLoad, kNWait, kNWaitVault, kNWaitArr, kNCounts };
struct WqWarpState
{
    int n[kNCounts];                // slots per state; kNWait / kNWaitVault: input / vault tickets held that could not be redeemed at the last LOAD
    unsigned retired;               // histories ended, not yet subtracted from the global in-flight count
    unsigned backoff;
    unsigned has_pub;               // some lane wrote fission secondaries in the last collision batch (pub_first / pub_n): publish them
    unsigned input_left;            // streamed input records may still be unclaimed (cleared once the input queue is seen empty)
    unsigned dry;                   // > 0: the last LOAD found nothing for some empty slots; counts down the batches until the next try
    unsigned pad;
    unsigned tally[12];             // the warp's balance counters (kTal*), flushed once at kernel end
    unsigned long long in_seen;     // host-buffer streaming: last value of ctl->in_ready the warp has seen
    unsigned long long t_start;
    long long c_start;              // clock64() at kernel start (the service-phase watchdog)
};

struct WqShared : SlotStore<kWqSlots>
{
    unsigned char state[kWqSlots];                  // what each slot waits for (kSt*)
    unsigned char list[kWqWarps][kWq];              // the warp's gather list: slots (warp-relative) of the batch being formed
    WqWarpState w[kWqWarps];
    unsigned long long pub_first[kWqThreads];       // per thread: first vault slot / number of the secondaries it wrote last
    unsigned pub_n[kWqThreads];
    PeerLaunch launch[kMaxPeers];
};

// gather the warp's slots whose state satisfies `pred` into its list (slot order); returns how many
template <class Pred>
__device__ __forceinline__ unsigned wq_gather(WqShared& s, unsigned warp, unsigned lane, Pred pred)
{
    unsigned total = 0;
#pragma unroll
    for (int k = 0; k < kWqK; ++k)
    {
        const bool hit = (kWq % 32 == 0 || 32 * k + (int)lane < kWq) && pred(s.state[warp * kWq + 32 * k + lane]);
        const unsigned m = __ballot_sync(kFullMask, hit);
        if (hit) s.list[warp][total + __popc(m & ((1u << lane) - 1u))] = (unsigned char)(32 * k + lane);
        total += __popc(m);
    }
    __syncwarp();
    return total;
}

// The events a history meets once (LOAD, CENSUS) or rarely (SEND) are separate functions: the hot loop -- choose, SEGMENT,
// COLLISION -- stays small enough for the instruction caches (the warp-state samples of the history kernel and of the first
// event kernels showed 10-17 % of cycles waiting for instructions).

// A fission secondary, raw (write_raw_child of track_physics.cuh): the parent's record as its slot holds it at the collision --
// the slot is only updated by the collision tail, after this -- with the child's stream and sampled outcome, slot -> vault a
// few fields at a time (see deposit_from_slot for why).
template <class Store>
__device__ __noinline__ void raw_child_from_slot(const TrackArgs& a, const Store& s, unsigned slot, unsigned long long i, uint64_t child_seed,
                                                 double energy_out, double angle_out)
{
    const VaultView& v = a.proc;
    __stcg(v.x + i, s.x[slot]); __stcg(v.y + i, s.y[slot]); __stcg(v.z + i, s.z[slot]); asm volatile("" ::: "memory");
#if QSB_VALIDATION
    __stcg(v.vx + i, s.vx[slot]); __stcg(v.vy + i, s.vy[slot]); __stcg(v.vz + i, s.vz[slot]); asm volatile("" ::: "memory");
#endif                          // fast build: the child's velocity is rebuilt by its collision tail before anything reads it
    __stcg(v.energy + i, energy_out); __stcg(v.weight + i, s.weight[slot]); __stcg(v.ttc + i, s.ttc[slot]); asm volatile("" ::: "memory");
    __stcg(v.age + i, s.age[slot]); __stcg(v.nmfp + i, angle_out); __stcg(v.nseg + i, s.nseg[slot]); asm volatile("" ::: "memory");
    __stcg(v.seed + i, (unsigned long long)child_seed); __stcg(v.id + i, (unsigned long long)child_seed);
    __stcg(v.cell + i, s.cell[slot]);
    __stcg(v.tags + i, make_int4(kRawChild, s.num_collisions[slot], s.breed[slot], s.species[slot])); asm volatile("" ::: "memory");
    __stcg(v.dirx + i, s.alpha[slot]); __stcg(v.diry + i, s.beta[slot]); __stcg(v.dirz + i, s.gamma[slot]);
}

// MC_Load_Particle (load_particle / load_particle_aos + reload_transform of track_physics.cuh: src/MC_Load_Particle.cc:11-29,
// src/MC_Base_Particle.hh:287-331), vault record -> slot, a few fields at a time with compiler barriers in between: like the
// census and the peer deposit this is a function whose register need, with the whole particle live, is paid for by spills in
// the tracking batches.  Same arithmetic on the same values, field for field.
template <class Store>
__device__ __forceinline__ int finish_loaded_slot(const TrackArgs& a, Store& s, unsigned slot, bool raw, bool derive_direction)
{
    // what the transform computes: |velocity|, the direction cosine (from the record, or re-derived from the velocity:
    // NaN marks "derive", the reference's MC_Particle(const MC_Base_Particle&) behaviour), the energy group
    double speed = 0.0;
    int group = 0;
    if (!raw)
    {
        const double vx = s.alpha[slot], vy = s.beta[slot], vz = s.gamma[slot];       // parked there by the caller when the direction is to be derived
        if (derive_direction)
        {
            speed = m_sqrt(vx * vx + vy * vy + vz * vz);
            const double factor = m_div(1.0, speed);
            s.alpha[slot] = factor * vx; s.beta[slot] = factor * vy; s.gamma[slot] = factor * vz;
        }
        group = energy_group(a, s.energy[slot]);
    }
    s.group[slot] = (unsigned short)group;
    s.facet[slot] = (unsigned char)0;
    if (!raw && derive_direction) s.speed[slot] = speed;
    return raw ? kStateTail : kStateSegment;
}

template <class Store>
__device__ __forceinline__ int load_slot_from_vault(const TrackArgs& a, Store& s, unsigned slot, unsigned long long i)
{
#define QSB_BARRIER asm volatile("" ::: "memory")
    const VaultView& v = a.proc;
    const int4 t = __ldcg(v.tags + i);
    const bool raw = t.x == kRawChild;                  // a fission secondary as its parent wrote it: the collision tail finishes it
    s.last_event[slot] = (unsigned char)(raw ? QSB_EV_COLLISION : t.x);
    s.num_collisions[slot] = t.y; s.breed[slot] = t.z; s.species[slot] = t.w;
    const int cell = __ldcg(v.cell + i);
    s.cell[slot] = cell;
    { const uint4 head = load_cell_head(a.im, cell); s.head01[slot] = make_uint2(head.x, head.y); s.events[slot] = head.z; } QSB_BARRIER;
    s.x[slot] = __ldcg(v.x + i); s.y[slot] = __ldcg(v.y + i); s.z[slot] = __ldcg(v.z + i); QSB_BARRIER;
    s.weight[slot] = __ldcg(v.weight + i); s.nmfp[slot] = __ldcg(v.nmfp + i); s.nseg[slot] = __ldcg(v.nseg + i); QSB_BARRIER;
    s.seed[slot] = __ldcg(v.seed + i); s.id[slot] = __ldcg(v.id + i); s.energy[slot] = __ldcg(v.energy + i); QSB_BARRIER;
    {
        double ttc = __ldcg(v.ttc + i), age = __ldcg(v.age + i);
        if (!raw) { if (ttc <= 0.0) ttc += a.dt; if (age < 0.0) age = 0.0; }
        s.ttc[slot] = ttc; s.age[slot] = age;
    }
    QSB_BARRIER;
    const double vx = __ldcg(v.vx + i), vy = __ldcg(v.vy + i), vz = __ldcg(v.vz + i);
#if QSB_VALIDATION
    s.vx[slot] = vx; s.vy[slot] = vy; s.vz[slot] = vz;
#endif
    const double dirx = __ldcg(v.dirx + i);
    const bool derive = !raw && dirx != dirx;
    if (derive) { s.alpha[slot] = vx; s.beta[slot] = vy; s.gamma[slot] = vz; }
    else
    {
        s.alpha[slot] = dirx; s.beta[slot] = __ldcg(v.diry + i); s.gamma[slot] = __ldcg(v.dirz + i);
        s.speed[slot] = raw ? 0.0 : m_sqrt(vx * vx + vy * vy + vz * vz);
    }
    QSB_BARRIER;
    return finish_loaded_slot(a, s, slot, raw, derive);
}

// the same from record i of the streamed host vault (136-byte MC_Base_Particle records, load_particle_aos)
template <class Store>
__device__ __forceinline__ int load_slot_from_records(const TrackArgs& a, Store& s, unsigned slot, unsigned long long i)
{
    const double* __restrict__ r = reinterpret_cast<const double*>(a.in_aos + i);
    const unsigned long long* __restrict__ u = reinterpret_cast<const unsigned long long*>(r);
    const unsigned long long t0 = __ldcg(u + 14), t1 = __ldcg(u + 15), t2 = __ldcg(u + 16);
    s.last_event[slot] = (unsigned char)(int)(unsigned)t0;
    s.num_collisions[slot] = (int)(unsigned)(t0 >> 32); s.breed[slot] = (int)(unsigned)t1; s.species[slot] = (int)(unsigned)(t1 >> 32);
    const int cell = __ldg(a.im.domain_cell_offset + (int)(unsigned)t2) + (int)(unsigned)(t2 >> 32);
    s.cell[slot] = cell;
    { const uint4 head = load_cell_head(a.im, cell); s.head01[slot] = make_uint2(head.x, head.y); s.events[slot] = head.z; } QSB_BARRIER;
    s.x[slot] = __ldcg(r + 0); s.y[slot] = __ldcg(r + 1); s.z[slot] = __ldcg(r + 2); QSB_BARRIER;
    s.weight[slot] = __ldcg(r + 7); s.nmfp[slot] = __ldcg(r + 10); s.nseg[slot] = __ldcg(r + 11); QSB_BARRIER;
    s.seed[slot] = __ldcg(u + 12); s.id[slot] = __ldcg(u + 13); s.energy[slot] = __ldcg(r + 6); QSB_BARRIER;
    {
        double ttc = __ldcg(r + 8), age = __ldcg(r + 9);
        if (ttc <= 0.0) ttc += a.dt;
        if (age < 0.0) age = 0.0;
        s.ttc[slot] = ttc; s.age[slot] = age;
    }
    QSB_BARRIER;
    const double vx = __ldcg(r + 3), vy = __ldcg(r + 4), vz = __ldcg(r + 5);
#if QSB_VALIDATION
    s.vx[slot] = vx; s.vy[slot] = vy; s.vz[slot] = vz;
#endif
    s.alpha[slot] = vx; s.beta[slot] = vy; s.gamma[slot] = vz;
    QSB_BARRIER;
    return finish_loaded_slot(a, s, slot, false, true);
#undef QSB_BARRIER
}

// LOAD: every slot of the warp without a particle; empty ones take a ticket (one global atomic per batch of 32), tickets
// are redeemed once the vault slot they name has been written -- the history kernel's protocol (its service phase).
//
// Two queues.  Tickets [0, n_in) are the records of a streamed host vault (host-buffer call: they land in HBM chunk by chunk
// while the kernel runs, PCIe-paced), tickets >= n_in are vault slots (initial population of a resident cycle, fission
// secondaries, arrivals from peers).  With ONE queue every secondary waits behind the whole input: once the kernel has caught
// up with the DMA front all slots hold input tickets, the secondaries pile up and are tracked -- and their census records
// sent back -- only after the last input chunk has landed (Coral2_P1: ~10 ms of a 38 ms call).  So the input has its own
// head (ctl->head_in), and while input is left a warp takes VAULT tickets only for slots that exist (below the tail) and
// input tickets for the rest of its empty slots; an input ticket is always redeemed eventually (the DMA front moves on its
// own), a vault ticket past the final tail never is -- harmless, the cycle ends on the in-flight count -- and the few of
// them a race can leave a warp with (two warps reading the same tail) are bounded by `may_take_vault`.
//
// Peer mode (arrival_cap != 0): a third queue, served first -- the arrival region other GPUs deposit into (device_types.cuh).
// There, too, vault tickets are only taken for slots that exist: a warp must keep empty slots for what the peers send it.
// `unserved` counts empty slots that found nothing to take; the scheduler then leaves the warp's refills alone for a few
// batches (WqWarpState::dry) instead of spinning on them -- an idle warp falls through to the termination test at once.
struct WqLoaded { int to_segment, to_tail, waiting_in, waiting_vault, waiting_arr, unserved; };
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// kPeer / kStream: the instance carries the arrival queue / the streamed-input queue; the resident single-GPU instance has
// neither (the tracking loop is instruction-cache bound: every instruction the refill path does not need is paid for in the
// SEGMENT and COLLISION batches, measured)
template <int kPeer, int kStream>
__device__ QSB_WQ_SERVICE_FN WqLoaded wq_load(const TrackArgs& a, WqShared& s, unsigned warp, unsigned lane, bool may_take_in, bool may_take_vault,
                                         bool may_take_arr)
{
    const bool has_arrivals = kPeer && a.arrival_cap != 0ull;
    const unsigned long long n_in = kStream ? a.n_in : 0ull;
    const unsigned base = warp * kWq;
    unsigned long long in_seen = s.w[warp].in_seen;
    const unsigned total = wq_gather(s, warp, lane, [](unsigned char st) { return st == kStLoad; });
    WqLoaded out = { 0, 0, 0, 0, 0, 0 };

    // ---- tickets for every empty slot of the warp, in ONE transaction: the queue counters are read together (one L2 round
    // trip), the heads are advanced together (a second one) -- a LOAD used to pay a dependent global atomic per batch of 32,
    // and three to five dependent accesses per batch once there were several queues.  Slot j of the list (batch j / 32, lane
    // j % 32) that has no ticket gets the next one; rank = its position among the wanting slots in list order.
    unsigned long long held[kWqK];
    unsigned before[kWqK];                              // wanting slots in earlier batches + lower lanes of this batch
    unsigned n_want = 0;
#pragma unroll
    for (int k = 0; k < kWqK; ++k)
    {
        const bool active = 32u * k + lane < total;
        held[k] = active ? s.id[base + s.list[warp][32 * k + lane]] : 0ull;       // 0: not a candidate (a real ticket of value 0 is not kNoTicket either)
        const unsigned m = __ballot_sync(kFullMask, active && held[k] == kNoTicket);
        before[k] = n_want + __popc(m & ((1u << lane) - 1u));
        n_want += __popc(m);
    }
    const bool has_prio = kPeer && QSB_OPT_BOUNDARY_FIRST && a.prio_list != nullptr;
    unsigned long long c_ptail = 0;
    if (n_want)
    {
        unsigned k_a = 0, k_v = 0, k_in = 0, k_p = 0;
        unsigned long long c_tail = 0, c_head = 0, c_atail = 0, c_ahead = 0, c_hin = 0, c_phead = 0;
        const bool input_maybe = kStream && n_in != 0ull && s.w[warp].input_left != 0u;
        if (has_arrivals || input_maybe)
        {
            // lane 0: vault tail, 1: vault head, 2: arrival tail, 3: arrival head, 4: input head, 5 / 6: boundary-first list head /
            // length -- one load instruction
            unsigned long long v = 0;
            const unsigned long long* src = lane == 0 ? a.tail : lane == 1 ? &a.ctl->head : lane == 2 ? &peer_control(a, a.my_rank)->arr_tail
                                          : lane == 3 ? &a.ctl->arr_head : lane == 4 ? &a.ctl->head_in : lane == 5 ? &a.ctl->prio_head : &a.ctl->prio_count;
            if (lane < 2 || (has_arrivals && lane < 4) || (input_maybe && lane == 4) || (has_prio && (lane == 5 || lane == 6))) v = ld_relaxed_u64(src);
            c_tail = __shfl_sync(kFullMask, v, 0); c_head = __shfl_sync(kFullMask, v, 1);
            if (kPeer) { c_atail = __shfl_sync(kFullMask, v, 2); c_ahead = __shfl_sync(kFullMask, v, 3); }
            if (kStream) c_hin = __shfl_sync(kFullMask, v, 4);
            if (kPeer) { c_phead = __shfl_sync(kFullMask, v, 5); c_ptail = __shfl_sync(kFullMask, v, 6); }
        }
        const bool input_left = input_maybe && c_hin < n_in;
        if (kStream && lane == 0 && !input_left) s.w[warp].input_left = 0u;
        unsigned left = n_want;
        if (has_arrivals && may_take_arr && c_atail > c_ahead)     // peer mode: particles other GPUs have deposited come first
        { k_a = (unsigned)min((unsigned long long)left, c_atail - c_ahead); left -= k_a; }
        if (has_prio && c_ptail > c_phead)                         // then the histories that may reach another GPU
        { k_p = (unsigned)min((unsigned long long)left, c_ptail - c_phead); left -= k_p; }
        if (input_left || has_arrivals)                            // vault tickets only for slots that exist
        {
            if (may_take_vault && c_tail > c_head) { k_v = (unsigned)min((unsigned long long)left, c_tail - c_head); left -= k_v; }
            if (input_left && may_take_in) k_in = left;
        }
        else k_v = may_take_vault ? left : 0u;
        if (kPeer && has_arrivals && lane == 0 && c_tail <= c_head)       // diagnostics: when did this GPU run out of its own work?
        {
            PeerControl* me = peer_control(a, a.my_rank);
            if (*((volatile unsigned long long*)&me->bulk_done_ns) == 0ull) me->bulk_done_ns = global_timer_ns() - s.w[warp].t_start;
        }
        // lane 0 advances the vault head, lane 1 the arrival head, lane 2 the input head -- one atomic instruction
        unsigned long long tb = 0;
        if (lane == 0 && k_v) tb = atomicAdd(&a.ctl->head, (unsigned long long)k_v);
        if (kPeer && lane == 1 && k_a) tb = atomicAdd(&a.ctl->arr_head, (unsigned long long)k_a);
        if (kStream && lane == 2 && k_in) tb = atomicAdd(&a.ctl->head_in, (unsigned long long)k_in);
        if (kPeer && lane == 3 && k_p) tb = atomicAdd(&a.ctl->prio_head, (unsigned long long)k_p);
        const unsigned long long tb_v = __shfl_sync(kFullMask, tb, 0);
        const unsigned long long tb_a = kPeer ? __shfl_sync(kFullMask, tb, 1) : 0ull;
        const unsigned long long tb_in = kStream ? __shfl_sync(kFullMask, tb, 2) : 0ull;
        const unsigned long long tb_p = kPeer ? __shfl_sync(kFullMask, tb, 3) : 0ull;
        out.unserved = (int)(n_want - k_a - k_p - k_v - k_in);
#pragma unroll
        for (int k = 0; k < kWqK; ++k)
        {
            if (held[k] != kNoTicket) continue;
            unsigned r = before[k];
            if (r < k_a) { held[k] = kArrivalTicket + tb_a + r; continue; }
            r -= k_a;
            if (r < k_p) { held[k] = kPrioTicket + tb_p + r; continue; }
            r -= k_p;
            if (r < k_v) { held[k] = tb_v + r; continue; }
            r -= k_v;
            if (kStream && r < k_in && tb_in + r < n_in) held[k] = tb_in + r;      // past the input's end: no ticket, next time a vault one
        }
    }

#pragma unroll 1
    for (unsigned first = 0, k = 0; first < total; first += 32u, ++k)      // (not unrolled: one copy of the particle-load code)
    {
        const bool active = first + lane < total;
        const unsigned slot = base + (active ? s.list[warp][first + lane] : 0u);
        unsigned long long mine = held[0];
#pragma unroll
        for (int j = 1; j < kWqK; ++j) if (k == (unsigned)j) mine = held[j];
        unsigned long long ticket = active ? mine : kNoTicket;
        bool ready = false;
        unsigned long long vslot_of = 0;                // vault slot of a non-input ticket
        if (active && ticket != kNoTicket)
        {
            if (kPeer && ticket >= kPrioTicket && ticket < kArrivalTicket)
            {
                // an entry of the boundary-first list: a slot of the initial population, ready by construction (an index past the
                // list's end can only come from two warps racing for its last entries: no ticket)
                const unsigned long long t = ticket - kPrioTicket;
                if (t < c_ptail) { vslot_of = __ldg(a.prio_list + t); ready = true; }
                else ticket = kNoTicket;
            }
            else if (kPeer && ticket >= kArrivalTicket)
            {
                const unsigned long long t = ticket - kArrivalTicket;
                if (t < a.arrival_cap)
                {
                    vslot_of = a.arrival_first + t;
                    uint32_t flag;
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(flag) : "l"(a.proc.ready + vslot_of) : "memory");
                    ready = flag == (a.epoch | kArrivalBit) && deposit_complete(a.proc, vslot_of, a.epoch);
                }
            }
            else if (kStream && ticket < n_in)
            {
                if (ticket >= in_seen)
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(in_seen) : "l"(&a.ctl->in_ready) : "memory");
                ready = ticket < in_seen;
            }
            else if (ticket - n_in < a.proc.capacity - (kPeer ? a.arrival_cap : 0ull))
            {
                const unsigned long long vslot = ticket - n_in;
                vslot_of = vslot;
                ready = vslot < a.ready_prefix;
                if (has_prio && vslot < a.prio_slots && __ldg(a.im.cell_near + __ldcg(a.proc.cell + vslot)) != 0)
                { ready = false; ticket = kNoTicket; }     // tracked through the boundary-first list: not this queue's
                else if (!ready)
                {
                    uint32_t flag;
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(flag) : "l"(a.proc.ready + vslot) : "memory");
                    ready = flag == a.epoch;
                    if (kPeer && !QSB_OPT_ARRIVAL_QUEUE && __builtin_expect(flag == (a.epoch | kArrivalBit), 0)) ready = deposit_complete(a.proc, vslot, a.epoch);
                }
            }
        }
        int state = kStateIdle;
        if (ready)
        {
            if (kStream && ticket < n_in) state = load_slot_from_records(a, s, slot, ticket);
            else state = load_slot_from_vault(a, s, slot, vslot_of);
            s.state[slot] = (unsigned char)(state == kStateTail ? kStTail : kStSegment);
        }
        else if (active) s.id[slot] = ticket;
        out.to_segment += __popc(__ballot_sync(kFullMask, state == kStateSegment));
        out.to_tail += __popc(__ballot_sync(kFullMask, state == kStateTail));
        if (kStream) out.waiting_in += __popc(__ballot_sync(kFullMask, active && !ready && ticket < n_in));
        out.waiting_vault += __popc(__ballot_sync(kFullMask, active && !ready && ticket >= n_in && ticket < kArrivalTicket));
        if (has_arrivals) out.waiting_arr += __popc(__ballot_sync(kFullMask, active && !ready && ticket >= kArrivalTicket && ticket != kNoTicket));
    }
    if (kStream)            // keep the highest DMA front any lane has seen
    {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { const unsigned long long o = __shfl_xor_sync(kFullMask, in_seen, d); in_seen = o > in_seen ? o : in_seen; }
        if (lane == 0) s.w[warp].in_seen = in_seen;
    }
    return out;
}

// CENSUS: up to 32 histories that reached census; their records go to the census vault (or the streamed record buffer).
// Written slot -> global memory one field at a time (compiler barrier after each), like deposit_from_slot and for the same
// reason: with the whole particle in registers this function's need, added to the kernel's at the call site, made the
// 128-register build spill in the SEGMENT and COLLISION batches.
template <int kStream, class Store>
__device__ __forceinline__ void census_from_slot(const TrackArgs& a, const Store& s, unsigned slot, unsigned long long i)
{
#define QSB_BARRIER asm volatile("" ::: "memory")
    const int cell = s.cell[slot];
    if (kStream && a.census_aos)        // 136-byte record, src/MC_Base_Particle.hh:75-92 (store_census_aos)
    {
        double* r = reinterpret_cast<double*>(a.census_aos + i);
        __stcg(r + 0, s.x[slot]); __stcg(r + 1, s.y[slot]); __stcg(r + 2, s.z[slot]); QSB_BARRIER;
#if QSB_VALIDATION
        __stcg(r + 3, s.vx[slot]); __stcg(r + 4, s.vy[slot]); __stcg(r + 5, s.vz[slot]); QSB_BARRIER;
#else
        { const double speed = s.speed[slot]; __stcg(r + 3, speed * s.alpha[slot]); __stcg(r + 4, speed * s.beta[slot]); __stcg(r + 5, speed * s.gamma[slot]); } QSB_BARRIER;
#endif
        __stcg(r + 6, s.energy[slot]); __stcg(r + 7, s.weight[slot]); __stcg(r + 8, s.ttc[slot]); QSB_BARRIER;
        __stcg(r + 9, s.age[slot]); __stcg(r + 10, s.nmfp[slot]); __stcg(r + 11, s.nseg[slot]); QSB_BARRIER;
        unsigned long long* u = reinterpret_cast<unsigned long long*>(r);
        __stcg(u + 12, s.seed[slot]); __stcg(u + 13, s.id[slot]); QSB_BARRIER;
        const int d = flat_to_domain(a.im, cell);
        const int local = cell - __ldg(a.im.domain_cell_offset + d);
        __stcg(u + 14, (unsigned long long)(unsigned)s.last_event[slot] | ((unsigned long long)(unsigned)s.num_collisions[slot] << 32));
        __stcg(u + 15, (unsigned long long)(unsigned)s.breed[slot] | ((unsigned long long)(unsigned)s.species[slot] << 32));
        __stcg(u + 16, (unsigned long long)(unsigned)d | ((unsigned long long)(unsigned)local << 32));
    }
    else                                // SoA census vault (store_particle, direction left to be re-derived: NaN)
    {
        const VaultView& v = a.census;
        __stcg(v.x + i, s.x[slot]); __stcg(v.y + i, s.y[slot]); __stcg(v.z + i, s.z[slot]); QSB_BARRIER;
#if QSB_VALIDATION
        __stcg(v.vx + i, s.vx[slot]); __stcg(v.vy + i, s.vy[slot]); __stcg(v.vz + i, s.vz[slot]); QSB_BARRIER;
#else
        { const double speed = s.speed[slot]; __stcg(v.vx + i, speed * s.alpha[slot]); __stcg(v.vy + i, speed * s.beta[slot]); __stcg(v.vz + i, speed * s.gamma[slot]); } QSB_BARRIER;
#endif
        __stcg(v.energy + i, s.energy[slot]); __stcg(v.weight + i, s.weight[slot]); __stcg(v.ttc + i, s.ttc[slot]); QSB_BARRIER;
        __stcg(v.age + i, s.age[slot]); __stcg(v.nmfp + i, s.nmfp[slot]); __stcg(v.nseg + i, s.nseg[slot]); QSB_BARRIER;
        __stcg(v.seed + i, s.seed[slot]); __stcg(v.id + i, s.id[slot]); QSB_BARRIER;
        __stcg(v.cell + i, cell);
        __stcg(v.tags + i, make_int4((int)s.last_event[slot], s.num_collisions[slot], s.breed[slot], s.species[slot]));
        const double nan = __longlong_as_double(0x7ff8000000000000ll);
        __stcg(v.dirx + i, nan); __stcg(v.diry + i, nan); __stcg(v.dirz + i, nan);
    }
#undef QSB_BARRIER
}

// kStream: the instance that can stream its census to the host as records (the launcher picks it whenever a.census_aos is set)
template <int kStream>
__device__ QSB_WQ_SERVICE_FN int wq_census(const TrackArgs& a, WqShared& s, unsigned warp, unsigned lane)
{
    const unsigned base = warp * kWq;
    const unsigned total = wq_gather(s, warp, lane, [](unsigned char st) { return st == kStCensus; });
    const bool active = lane < total;                   // the batch: lanes 0 .. count-1, list order
    const unsigned slot = base + (active ? s.list[warp][lane] : 0u);
    const unsigned count = min(total, 32u);
    unsigned long long cbase = 0;
    if (lane == 0) cbase = atomicAdd(&a.ctl->census_count, (unsigned long long)count);
    cbase = __shfl_sync(kFullMask, cbase, 0);
    if (active)
    {
        if (cbase + lane >= a.census.capacity) atomicOr(&a.ctl->overflow, 2u);
        else census_from_slot<kStream>(a, s, slot, cbase + lane);
        s.id[slot] = kNoTicket; s.state[slot] = (unsigned char)kStLoad;
    }
    if (kStream && a.census_aos)
    {
        // streaming: count the batch's records into their chunk(s) with release semantics -- the records of the whole warp
        // (ordered by the __syncwarp) are visible before the count -- and tell the host about every chunk that became
        // complete, through mapped pinned memory; its D2H copy then runs while tracking continues (census_flush)
        __syncwarp();
        if (lane == 0)
        {
            const unsigned long long last = min(cbase + count, a.census.capacity);
            unsigned long long at = cbase;
            while (at < last)
            {
                const unsigned long long chunk = at >> a.census_chunk_shift;
                const unsigned long long chunk_end = min((chunk + 1) << a.census_chunk_shift, last);
                const unsigned cnt = (unsigned)(chunk_end - at);
                unsigned old;
                asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(a.census_chunk_done + chunk), "r"(cnt) : "memory");
                if (old + cnt == (1u << a.census_chunk_shift))
                    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(a.host_chunk_flags + chunk), "r"(a.epoch) : "memory");
                at = chunk_end;
            }
        }
    }
    return (int)count;
}

// SEND: up to 32 particles whose facet crossing leaves this rank's domain.  NCCL-rounds mode: packed into the per-peer slab
// (facet_crossing_event's own code path).  Peer mode: deposited straight into the neighbour GPU's processing vault over
// NVLink; the remote counters are raised once per (batch, destination) instead of once per particle, in the order the
// termination test relies on (peer's `sent` and `inflight` before the record is stored and the history retired here; peer's
// `received` after its `inflight`), see send_advance in track_physics.cuh.
// A particle's record, straight from its slot into slot i of a PEER's vault (what store_deposit does from registers): one
// field at a time, with a compiler barrier after each, so that the function needs a handful of registers instead of the
// whole particle.  That matters far from here: the register need of this rarely-run function is added to what the kernel
// keeps live at its call site, and the allocator answered by spilling in the SEGMENT and COLLISION batches of the
// peer-exchange instance (measured: +14 % at 168 registers, +36 % at 128).
template <class Store>
__device__ __noinline__ void deposit_from_slot(const Store& s, unsigned slot, char* peer_base, unsigned long long cap, unsigned long long i,
                                               int cell, uint32_t vault_epoch)
{
    // array k of vault_view starts at d + k * cap; a RUNNING pointer walks from array to array (twenty precomputed addresses
    // would be forty registers, and the allocator would take them from the tracking batches)
    double* d = reinterpret_cast<double*>(peer_base + kVaultHeaderBytes);
    volatile unsigned long long step = cap;             // opaque to the optimiser: keeps the address arithmetic sequential
    double* ptr = d + i;
    unsigned long long x = deposit_salt(vault_epoch) ^ (unsigned long long)(unsigned)cell;
#define QSB_DEPOSIT_F64(value_) { const double v_ = (value_); __stcg(ptr, v_); x ^= bits(v_); ptr += step; asm volatile("" ::: "memory"); }
    QSB_DEPOSIT_F64(s.x[slot]) QSB_DEPOSIT_F64(s.y[slot]) QSB_DEPOSIT_F64(s.z[slot])
#if QSB_VALIDATION
    QSB_DEPOSIT_F64(s.vx[slot]) QSB_DEPOSIT_F64(s.vy[slot]) QSB_DEPOSIT_F64(s.vz[slot])
#else
    QSB_DEPOSIT_F64(s.speed[slot] * s.alpha[slot]) QSB_DEPOSIT_F64(s.speed[slot] * s.beta[slot]) QSB_DEPOSIT_F64(s.speed[slot] * s.gamma[slot])
#endif
    QSB_DEPOSIT_F64(s.energy[slot]) QSB_DEPOSIT_F64(s.weight[slot]) QSB_DEPOSIT_F64(s.ttc[slot]) QSB_DEPOSIT_F64(s.age[slot])
    QSB_DEPOSIT_F64(s.nmfp[slot]) QSB_DEPOSIT_F64(s.nseg[slot])
    QSB_DEPOSIT_F64(s.alpha[slot]) QSB_DEPOSIT_F64(s.beta[slot]) QSB_DEPOSIT_F64(s.gamma[slot])
#undef QSB_DEPOSIT_F64
    unsigned long long* u = reinterpret_cast<unsigned long long*>(ptr);                 // array 15: seed
    { const unsigned long long v = s.seed[slot]; __stcg(u, v); x ^= v; u += step; }    // 16: id
    { const unsigned long long v = s.id[slot];   __stcg(u, v); x ^= v; u += step; }    // 17: check (stored last)
    asm volatile("" ::: "memory");
    const int4 tags = make_int4((int)s.last_event[slot], s.num_collisions[slot], s.breed[slot], s.species[slot]);
    x ^= (unsigned long long)(unsigned)tags.x | ((unsigned long long)(unsigned)tags.y << 32);
    x ^= (unsigned long long)(unsigned)tags.z | ((unsigned long long)(unsigned)tags.w << 32);
    __stcg(reinterpret_cast<int4*>(d + 18 * step) + i, tags);
    int* cells = reinterpret_cast<int*>(d + 20 * step);
    __stcg(cells + i, cell);
    __stcg(u, x);
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(reinterpret_cast<uint32_t*>(cells + step) + i), "r"(vault_epoch | kArrivalBit) : "memory");
}

template <int kPeer>
__device__ QSB_WQ_SERVICE_FN int wq_send(const TrackArgs& a, WqShared& s, unsigned warp, unsigned lane)
{
    const long long t_enter = clock64();
    const unsigned base = warp * kWq;
    const unsigned total = wq_gather(s, warp, lane, [](unsigned char st) { return st == kStSend; });
    const bool active = lane < total;
    const unsigned slot = base + (active ? s.list[warp][lane] : 0u);
    if (!(kPeer && a.peer_mode))
    {
        Particle p;
        Counters unused = {};
        if (active) { load_all(s, slot, p); facet_crossing_event(a, p, unused); }     // TRANSIT_OFF: writes the exchange record into the per-peer slab
    }
    else if (kPeer)
    {
        const DevImage& im = a.im;
        int rank = -1;
        size_t k = 0;
        if (active) { k = (size_t)s.cell[slot] * 6 + (s.facet[slot] >> 2); rank = __ldg(im.face_nbr_rank + k); }
        unsigned todo = __ballot_sync(kFullMask, active);
        while (todo)
        {
            const unsigned head_lane = __ffs(todo) - 1;
            const int dest = __shfl_sync(kFullMask, rank, head_lane);
            const unsigned group = __ballot_sync(kFullMask, active && rank == dest);
            const unsigned n = __popc(group);
            unsigned long long ticket0 = 0, dep = 0, dep2 = 0;
            PeerControl* pc = peer_control(a, dest);
            if (lane == head_lane)
            {
                atomicAdd(&a.ctl->send_count[dest], (unsigned long long)n);          // statistics only
                dep = atomicAdd_system(&pc->sent, (unsigned long long)n);
                dep2 = atomicAdd_system(&pc->inflight, (unsigned long long)n);
#if QSB_OPT_ARRIVAL_QUEUE
                ticket0 = atomicAdd_system(&pc->arr_tail, (unsigned long long)n);
#else
                ticket0 = atomicAdd_system(&pc->tail, (unsigned long long)n);
#endif
                asm volatile("" :: "l"(dep), "l"(dep2) : "memory");                 // performed: their values are here
            }
            ticket0 = __shfl_sync(kFullMask, ticket0, head_lane);
            if (active && rank == dest)
            {
#if QSB_OPT_ARRIVAL_QUEUE
                const unsigned long long t_arr = ticket0 + __popc(group & ((1u << lane) - 1u));
                const unsigned long long vslot = t_arr < a.arrival_cap ? a.arrival_first + t_arr : a.proc.capacity;
#else
                const unsigned long long vslot = ticket0 + __popc(group & ((1u << lane) - 1u)) - s.launch[dest].n_in;
#endif
                if (vslot >= a.proc.capacity)
                {
                    st_release_sys(&pc->overflow, a.peer_epoch);                     // the peer's host reports it; the particle is dropped
                    atomicAdd_system(&pc->inflight, 0ull - 1ull);
                }
                else
                    deposit_from_slot(s, slot, a.peer_base[dest], a.proc.capacity, vslot, peer_destination_cell(a, pc, k), s.launch[dest].vault_epoch);
            }
            __syncwarp();
            if (lane == head_lane)
                asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" :: "l"(&pc->received), "l"((unsigned long long)n) : "memory");
            todo &= ~group;
        }
    }
    if (active) { s.id[slot] = kNoTicket; s.state[slot] = (unsigned char)kStLoad; }
    if (kPeer && a.peer_mode && lane == 0)
    {
        atomicAdd(&peer_control(a, a.my_rank)->send_cycles, (unsigned long long)(clock64() - t_enter));
        atomicAdd(&peer_control(a, a.my_rank)->send_calls, 1ull);
    }
    return (int)min(total, 32u);
}

template <int kDummy, int kPeer, int kStream>
__global__ void __launch_bounds__(kWqThreads, QSB_WQ_MIN_BLOCKS) track_warpq_kernel(const __grid_constant__ TrackArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WqShared& s = *reinterpret_cast<WqShared*>(smem_raw);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned base = warp * kWq;

    if (kPeer && a.peer_mode)
    {
        if (blockIdx.x == 0)                                    // the service block: global termination, tracks nothing
        {
            if (threadIdx.x < 32u)
            {
                const PeerControl* pc = peer_control(a, (int)min(threadIdx.x, (unsigned)(a.im.n_ranks - 1)));
                const unsigned long long t0 = global_timer_ns();
                while (ld_acquire_sys(&pc->epoch) != a.peer_epoch && global_timer_ns() - t0 < a.watchdog_ns) __nanosleep(500);
                if (threadIdx.x == 0) peer_control(a, a.my_rank)->startup_wait_ns = global_timer_ns() - t0;
                __syncwarp();
                peer_service_loop(a, lane);
            }
            return;
        }
        if ((int)threadIdx.x < a.im.n_ranks)
        {
            const PeerControl* pc = peer_control(a, (int)threadIdx.x);
            const unsigned long long t0 = global_timer_ns();
            while (ld_acquire_sys(&pc->epoch) != a.peer_epoch && global_timer_ns() - t0 < a.watchdog_ns) __nanosleep(500);
            s.launch[threadIdx.x].n_in = ld_relaxed_sys(&pc->n_in);
            s.launch[threadIdx.x].vault_epoch = ld_relaxed_sys(&pc->vault_epoch);
        }
        __syncthreads();
    }

#pragma unroll
    for (int k = 0; k < kWqK; ++k)
        if (kWq % 32 == 0 || 32 * k + (int)lane < kWq) { s.state[base + 32 * k + lane] = (unsigned char)kStLoad; s.id[base + 32 * k + lane] = kNoTicket; }
    s.pub_n[threadIdx.x] = 0u;
    if (lane == 0)
    {
        WqWarpState& w0 = s.w[warp];
        for (int k = 0; k < kNCounts; ++k) w0.n[k] = 0;
        w0.n[kNLoad] = kWq;
        w0.retired = 0u; w0.backoff = 64u; w0.has_pub = 0u; w0.input_left = 1u; w0.dry = 0u; w0.pad = 0u;
        for (int k = 0; k < 12; ++k) w0.tally[k] = 0u;
        w0.in_seen = 0ull; w0.t_start = global_timer_ns(); w0.c_start = clock64();
    }

#ifdef QSB_EXP_BLOAT
    // experiment: ~2000 instructions that never run (a.dt is positive): does the sheer size of the kernel cost time?
    if (a.dt < -1.0)
    {
        double v = a.dt;
#pragma unroll
        for (int k = 0; k < QSB_EXP_BLOAT; ++k) { v = qs_strict_log(v * v + 1.5) + (double)k; if (v > 3.0) __nanosleep(10); }
        if (v == 0.123) atomicOr(&a.ctl->overflow, 16u);
    }
#endif
    for (;;)
    {
        __syncwarp();
        WqWarpState& w = s.w[warp];
        const int n_seg = w.n[kNSeg], n_col = w.n[kNCol], n_cen = w.n[kNCen], n_snd = w.n[kNSnd], n_load = w.n[kNLoad];
        const int n_wait_in = w.n[kNWait], n_wait_vault = w.n[kNWaitVault], n_wait_arr = w.n[kNWaitArr], n_wait = n_wait_in + n_wait_vault + n_wait_arr;
        const unsigned has_pub = w.has_pub, w_input_left = w.input_left, w_dry = w.dry;
        __syncwarp();                                   // everybody has read the state before lane 0 may change it
        if (w_dry && lane == 0) w.dry = w_dry - 1u;
        if (has_pub)
        {
            const unsigned pub_n = s.pub_n[threadIdx.x];
            publish_children(a, s.pub_first[threadIdx.x], pub_n, a.epoch);
            s.pub_n[threadIdx.x] = 0u;
            if (lane == 0) w.has_pub = 0u;
        }

        // ---- which event next ----
        // Slots whose history has ended (census record to store, particle to ship to a peer) or that are empty do not take
        // part in the tracking batches: they are "parked".  Once kService of them have gathered -- or when no full tracking
        // batch is left and they outnumber what is -- the warp runs its SERVICE phase: every pending send, every pending
        // census record, then a refill of all empty slots, in one go (a particle bound for a peer must not sit in its slot
        // until 31 others have joined it: at ~140 sends per warp and cycle that parks a sixth of the warp's slots for good;
        // measured on 2 GPUs).  Otherwise: full collision batches (the longest event) before full segment batches.
        // tickets of a queue are handed out in order: holding unredeemable ones means its end (DMA front / tail) is reached
        // (n_fill: empty slots a LOAD can be expected to give a ticket to -- input tickets while input is left, else vault
        // tickets; once the warp holds its quota of unredeemable ones it counts as idle and looks at the termination test)
        const bool may_take_in = n_wait_in < 32, may_take_vault = n_wait_vault < ((kStream && a.n_in) ? 16 : 32), may_take_arr = n_wait_arr < 32;
        const bool input_left = kStream && a.n_in != 0ull && w_input_left != 0u;
        const int n_fill = ((input_left ? may_take_in : may_take_vault) && w_dry == 0u) ? n_load - n_wait : 0;
        const int n_parked = n_cen + n_snd + n_fill;
        int type;
#if QSB_OPT_SERVICE_ALL
        bool do_census = true, do_load = true;
        if (n_parked >= kService) type = kStLoad;
#else
        // which of the services: the one with most slots waiting (sends ride along with either, see below)
        bool do_census = n_cen + n_snd >= n_fill, do_load = !do_census;
        if (n_parked >= kService) type = kStLoad;
#endif
        else if (n_col >= 32) type = kStCollision;
        else if (n_seg >= 32) type = kStSegment;
        else
        {
            type = kStSegment; int best = n_seg;
            if (n_col > best) { best = n_col; type = kStCollision; }
#if QSB_OPT_SERVICE_ALL
            if (n_parked > best) { best = n_parked; type = kStLoad; }
#else
            if (n_fill > best) { best = n_fill; type = kStLoad; do_load = true; do_census = false; }
            if (n_cen + n_snd > best) { best = n_cen + n_snd; type = kStLoad; do_census = true; do_load = false; }
#endif
            if (best == 0)
            {
                // only tickets that cannot be redeemed yet: the warp is idle.  The cycle is over when no history is queued or
                // running anywhere on this GPU -- or, in peer mode, anywhere on any GPU (the service warp's verdict)
                unsigned done = 0u, backoff = 64u;
                if (lane == 0)
                {
                    if (w.retired) { atomicAdd(a.inflight, 0ull - (unsigned long long)w.retired); w.retired = 0u; }
                    if (!(kPeer && a.peer_mode))
                    {
                        done = atomicAdd(a.inflight, 0ull) == 0ull ? 1u : 0u;
                        if (!done && global_timer_ns() - w.t_start > (a.watchdog_ns ? a.watchdog_ns : 20000000000ull))
                        { atomicOr(&a.ctl->overflow, 8u); done = 1u; }       // never hang the device: give up, loudly
                    }
                    else
                    {
                        const PeerControl* me = peer_control(a, a.my_rank);
                        const unsigned d = *((volatile const unsigned int*)&me->done), ab = *((volatile const unsigned int*)&me->abort);
                        done = (d == a.peer_epoch || ab == a.peer_epoch) ? 1u : 0u;
                    }
                    backoff = w.backoff;
                    if (backoff < 2048u) w.backoff = backoff * 2u;
                }
                if (__shfl_sync(kFullMask, done, 0)) break;
                __nanosleep(__shfl_sync(kFullMask, backoff, 0));
                type = kStLoad; do_load = true; do_census = false;
            }
        }

        if (type == kStLoad)            // the service phase: SEND, CENSUS, LOAD
        {
            // sends ride along with a service phase once a few have gathered or when the warp has no full batch to run anyway:
            // a deposit costs an NVLink round trip (~5 us, measured) whatever the number of particles in it
            for (int left = (n_snd >= 8 || (n_seg < 32 && n_col < 32)) ? n_snd : 0; left > 0; left -= 32)
            {
                const int n_act = wq_send<kPeer>(a, s, warp, lane);
                if (lane == 0) { w.n[kNSnd] -= n_act; w.n[kNLoad] += n_act; w.retired += (unsigned)n_act; }
            }
            for (int left = do_census ? n_cen : 0; left > 0; left -= 32)
            {
                const int n_act = wq_census<kStream>(a, s, warp, lane);
                if (lane == 0) { w.n[kNCen] -= n_act; w.n[kNLoad] += n_act; w.retired += (unsigned)n_act; }
            }
            if (!do_load) continue;
            unsigned expired = 0u;
            if (lane == 0)
            {
                if (w.retired) { atomicAdd(a.inflight, 0ull - (unsigned long long)w.retired); w.retired = 0u; }
                // never hang the device, whatever state a warp is in: the watchdog is looked at in every service phase (the SM's
                // cycle counter, ~2 per ns: reading %globaltimer here, a few hundred times per warp and cycle, is not free)
#if QSB_OPT_SERVICE_WATCHDOG
                if ((unsigned long long)(clock64() - w.c_start) > 2ull * (a.watchdog_ns ? a.watchdog_ns : 20000000000ull))
#else
                if (false)
#endif
                {
                    expired = 1u;
                    atomicOr(&a.ctl->overflow, 8u);         // (peer mode: the service warp's own watchdog tells the other ranks)
                }
            }
            if (__shfl_sync(kFullMask, expired, 0)) break;
            const WqLoaded got = wq_load<kPeer, kStream>(a, s, warp, lane, may_take_in, may_take_vault, may_take_arr);
            if (lane == 0)
            {
                w.n[kNSeg] += got.to_segment; w.n[kNCol] += got.to_tail; w.n[kNLoad] -= got.to_segment + got.to_tail;
                w.n[kNWait] = got.waiting_in; w.n[kNWaitVault] = got.waiting_vault; w.n[kNWaitArr] = got.waiting_arr;
                w.dry = got.unserved ? 8u : 0u;
                if (got.to_segment | got.to_tail) w.backoff = 64u;
            }
            continue;
        }

        if (type == kStSegment)
        {
            const unsigned total = wq_gather(s, warp, lane, [](unsigned char st) { return st == kStSegment; });
            const bool active = lane < total;
            const unsigned slot = base + (active ? s.list[warp][lane] : 0u);
            int next = -1;
#if QSB_VALIDATION
            unsigned slow = 0u, mismatch = 0u;
#endif
            if (active)
            {
                Counters c = {};
                Particle p;
                p.x = s.x[slot]; p.y = s.y[slot]; p.z = s.z[slot];
                p.alpha = s.alpha[slot]; p.beta = s.beta[slot]; p.gamma = s.gamma[slot];
                p.weight = s.weight[slot]; p.ttc = s.ttc[slot]; p.age = s.age[slot];
                p.nmfp = s.nmfp[slot]; p.nseg = s.nseg[slot]; p.speed = s.speed[slot];
#if QSB_VALIDATION
                p.vx = s.vx[slot]; p.vy = s.vy[slot]; p.vz = s.vz[slot];
#endif
                p.seed = s.seed[slot];
                load_head(s, slot, p);
                p.cell = s.cell[slot];
                p.group = s.group[slot];
                p.facet = 0; p.last_event = 0; p.species = 0; p.total_xs = 0.0;

                const int outcome = segment_outcome(a, p, c);
                p.nseg += 1.;
                if (outcome == 0) next = kStCollision;
                else if (outcome == 1)
                {
                    if (face_event(p.head, p.facet >> 2) == QSB_ADJ_TRANSIT_OFF)
                    {
                        p.last_event = QSB_EV_COMMUNICATION;        // leaves the rank's domain: shipped by its own event (SEND), which needs every field
                        next = kStSend;
                    }
                    else next = facet_crossing_event(a, p, c) == 1 ? kStSegment : kStLoad;      // kStLoad: escaped, the slot is free
                }
                else next = kStCensus;
#if QSB_VALIDATION
                slow = c.slow; mismatch = c.mismatch;
#endif

                s.x[slot] = p.x; s.y[slot] = p.y; s.z[slot] = p.z;
                s.ttc[slot] = p.ttc; s.age[slot] = p.age; s.nmfp[slot] = p.nmfp; s.nseg[slot] = p.nseg;
                s.seed[slot] = p.seed;
                s.last_event[slot] = (unsigned char)p.last_event;
                if (outcome == 1)
                {
                    s.facet[slot] = (unsigned char)p.facet;
                    s.cell[slot] = p.cell;
                    store_head(s, slot, p);
                    s.alpha[slot] = p.alpha; s.beta[slot] = p.beta; s.gamma[slot] = p.gamma;       // reflection
#if QSB_VALIDATION
                    s.vx[slot] = p.vx; s.vy[slot] = p.vy; s.vz[slot] = p.vz; s.speed[slot] = p.speed;
#endif
                }
                if (next == kStLoad) s.id[slot] = kNoTicket;
                s.state[slot] = (unsigned char)next;
            }
            // where the batch went: the warp's slot counts and its balance tallies (every active lane advanced one segment;
            // a history that left SEGMENT for LOAD escaped, src/MC_Facet_Crossing_Event.cc:41-47)
            const int to_col = __popc(__ballot_sync(kFullMask, next == kStCollision));
            const int to_cen = __popc(__ballot_sync(kFullMask, next == kStCensus));
            const int to_load = __popc(__ballot_sync(kFullMask, next == kStLoad));
            const int to_snd = (kPeer || a.im.n_ranks > 1) ? __popc(__ballot_sync(kFullMask, next == kStSend)) : 0;
#if QSB_VALIDATION
            const unsigned n_slow = __reduce_add_sync(kFullMask, slow), n_mis = __reduce_add_sync(kFullMask, mismatch);
#endif
            if (lane == 0)
            {
                w.n[kNSeg] -= to_col + to_cen + to_load + to_snd;        // the rest of the batch stays in SEGMENT
                w.n[kNCol] += to_col; w.n[kNCen] += to_cen; w.n[kNLoad] += to_load; w.n[kNSnd] += to_snd;
                w.retired += (unsigned)to_load;
                w.tally[kTalSegments] += min(total, 32u); w.tally[kTalCensus] += (unsigned)to_cen; w.tally[kTalEscapes] += (unsigned)to_load;
#if QSB_VALIDATION
                w.tally[kTalSlow] += n_slow; w.tally[kTalMismatch] += n_mis;
#endif
            }
            continue;
        }

        if (type == kStCollision)
        {
            const unsigned total = wq_gather(s, warp, lane, [](unsigned char st) { return st == kStCollision || st == kStTail; });
            const bool active = lane < total;
            const unsigned slot = base + (active ? s.list[warp][lane] : 0u);
            // Three phases, each with only the fields it needs in registers (the rest of the particle stays in its slot): reaction
            // and sampled outcomes (energy, random-number stream, cross sections) -- secondaries (rare; the parent's record is
            // copied from the slot) -- outgoing trajectory (direction, census clock).
            Counters c = {};
            bool tail = false;
            double energy0 = 0.0, angle0 = 0.0, energy1 = 0.0, angle1 = 0.0, energy2 = 0.0, angle2 = 0.0, energy3 = 0.0, angle3 = 0.0;
            int n_out = 0;
            uint64_t seed = 0;
            if (active)
            {
                Particle p;
                tail = s.state[slot] == kStTail;
                p.energy = s.energy[slot];
                p.seed = s.seed[slot];
                load_head(s, slot, p);
                p.group = s.group[slot];
                // the total cross section the segment ended on (MC_Segment_Outcome leaves it in the particle, src/MC_Segment_Outcome.cc:60-66)
                p.total_xs = __ldg(a.im.xs_pair + (size_t)cell_material(p.head) * a.im.n_groups + p.group).x;
                energy0 = p.energy; angle0 = s.nmfp[slot];              // a raw child carries its sampled outcome in these two fields
                n_out = 1;
                if (!tail) n_out = collision_head(a, p, c, energy0, angle0, energy1, angle1, energy2, angle2, energy3, angle3);
                seed = p.seed;
            }
            // balance tallies of the batch (src/CollisionEvent.cc:104-119), from ballots: nothing per-thread stays live
            const unsigned t_col = __popc(__ballot_sync(kFullMask, c.collisions != 0u));
            const unsigned t_abs = __popc(__ballot_sync(kFullMask, c.absorbs != 0u));
            const unsigned fis_mask = __ballot_sync(kFullMask, c.fissions != 0u);
            const unsigned t_pro = fis_mask ? __reduce_add_sync(kFullMask, c.produced) : 0u;
#if QSB_VALIDATION
            const unsigned t_sca = __popc(__ballot_sync(kFullMask, c.scatters != 0u));
            const unsigned t_look = __reduce_add_sync(kFullMask, c.lookups), t_mis = __reduce_add_sync(kFullMask, c.mismatch);
#endif
            // secondaries of the whole batch: vault slots and the in-flight count with one atomic each
            const unsigned n_child = (active && !tail && n_out > 1) ? (unsigned)(n_out - 1) : 0u;
            unsigned incl = n_child;
            if (fis_mask)
            {
#pragma unroll
                for (int d = 1; d < 32; d <<= 1)
                {
                    const unsigned up = __shfl_up_sync(kFullMask, incl, d);
                    if ((int)lane >= d) incl += up;
                }
            }
            const unsigned n_children = fis_mask ? __shfl_sync(kFullMask, incl, 31) : 0u;
            if (n_children)
            {
                unsigned long long cb = 0;
                if (lane == 0)
                {
                    cb = atomicAdd(a.tail, (unsigned long long)n_children);
                    atomicAdd(a.inflight, (unsigned long long)n_children);
                }
                cb = __shfl_sync(kFullMask, cb, 0) - a.n_in;
                if (n_child)
                {
                    const unsigned long long first = cb + (incl - n_child);
                    if (first + n_child > a.proc.capacity - a.arrival_cap)
                    {
                        atomicOr(&a.ctl->overflow, 1u);
                        atomicAdd(a.inflight, 0ull - (unsigned long long)n_child);
                    }
                    else
                    {
                        raw_child_from_slot(a, s, slot, first, qs_rng_spawn(&seed), energy1, angle1);
                        if (n_child > 1u) raw_child_from_slot(a, s, slot, first + 1, qs_rng_spawn(&seed), energy2, angle2);
                        if (n_child > 2u) raw_child_from_slot(a, s, slot, first + 2, qs_rng_spawn(&seed), energy3, angle3);
                        s.pub_first[threadIdx.x] = first; s.pub_n[threadIdx.x] = n_child;       // published at the top of the next iteration
                    }
                }
            }
            bool absorbed = false;
            if (active)
            {
                if (n_out > 0)
                {
                    Particle p;
                    p.alpha = s.alpha[slot]; p.beta = s.beta[slot]; p.gamma = s.gamma[slot];
                    p.ttc = s.ttc[slot]; p.age = s.age[slot];
                    p.seed = seed;
                    collision_tail(a, p, energy0, angle0, tail || n_out > 1);
                    s.energy[slot] = p.energy;
                    s.alpha[slot] = p.alpha; s.beta[slot] = p.beta; s.gamma[slot] = p.gamma;
                    s.speed[slot] = p.speed; s.nmfp[slot] = p.nmfp; s.ttc[slot] = p.ttc; s.age[slot] = p.age;
#if QSB_VALIDATION
                    s.vx[slot] = p.vx; s.vy[slot] = p.vy; s.vz[slot] = p.vz;
#endif
                    s.seed[slot] = p.seed;
                    s.group[slot] = (unsigned short)p.group;
                    s.last_event[slot] = (unsigned char)QSB_EV_COLLISION;
                    s.state[slot] = (unsigned char)kStSegment;
                }
                else { absorbed = true; s.id[slot] = kNoTicket; s.state[slot] = (unsigned char)kStLoad; }
            }
            const int n_act = (int)min(total, 32u);
            const int n_gone = __popc(__ballot_sync(kFullMask, absorbed));
            if (lane == 0)
            {
                w.n[kNCol] -= n_act; w.n[kNSeg] += n_act - n_gone; w.n[kNLoad] += n_gone;
                w.retired += (unsigned)n_gone;
                if (n_children) w.has_pub = 1u;
                w.tally[kTalCollisions] += t_col; w.tally[kTalAbsorbs] += t_abs;
                if (fis_mask) { w.tally[kTalFissions] += (unsigned)__popc(fis_mask); w.tally[kTalProduced] += t_pro; }
#if QSB_VALIDATION
                w.tally[kTalScatters] += t_sca; w.tally[kTalLookups] += t_look; w.tally[kTalMismatch] += t_mis;
#endif
            }
            continue;
        }

    }

    // flush the warp's balance counters: one atomic per counter
    __syncwarp();
    if (lane == 0)
    {
        const unsigned* t = s.w[warp].tally;
        unsigned long long* b = a.ctl->balance;
#if QSB_VALIDATION
        const unsigned s_sca = t[kTalScatters];
#else
        // every collision of the fast build is one of the three reaction types (an undefined type would have been rejected by the host model)
        const unsigned s_sca = t[kTalCollisions] - t[kTalAbsorbs] - t[kTalFissions];
#endif
        if (t[kTalSegments]) atomicAdd(b + QSB_BAL_NUM_SEGMENTS, (unsigned long long)t[kTalSegments]);
        if (t[kTalCollisions]) atomicAdd(b + QSB_BAL_COLLISION, (unsigned long long)t[kTalCollisions]);
        if (s_sca) atomicAdd(b + QSB_BAL_SCATTER, (unsigned long long)s_sca);
        if (t[kTalAbsorbs]) atomicAdd(b + QSB_BAL_ABSORB, (unsigned long long)t[kTalAbsorbs]);
        if (t[kTalFissions]) atomicAdd(b + QSB_BAL_FISSION, (unsigned long long)t[kTalFissions]);
        if (t[kTalProduced]) atomicAdd(b + QSB_BAL_PRODUCE, (unsigned long long)t[kTalProduced]);
        if (t[kTalEscapes]) atomicAdd(b + QSB_BAL_ESCAPE, (unsigned long long)t[kTalEscapes]);
        if (t[kTalCensus]) atomicAdd(b + QSB_BAL_CENSUS, (unsigned long long)t[kTalCensus]);
        if (t[kTalLookups]) atomicAdd(&a.ctl->n_lookups, (unsigned long long)t[kTalLookups]);
        if (t[kTalSlow]) atomicAdd(&a.ctl->slow_geometry, (unsigned long long)t[kTalSlow]);
        if (t[kTalMismatch]) atomicAdd(&a.ctl->geometry_mismatch, (unsigned long long)t[kTalMismatch]);
    }
}

} // namespace

#if QSB_VALIDATION
#define QSB_EVT_LAUNCH_NAME launch_track_event_validation
#define QSB_EVT_ATTR_NAME track_event_kernel_attributes_validation
#else
#define QSB_EVT_LAUNCH_NAME launch_track_event_fast
#define QSB_EVT_ATTR_NAME track_event_kernel_attributes_fast
#endif

void QSB_EVT_LAUNCH_NAME(const TrackArgs& a, int grid, cudaStream_t s)
{
    // four instances: with / without the peer-exchange code, with / without the streamed-input queue.
    // QSB_FORCE_PEER_INSTANCE=1 (measurements only): run the instance that carries the peer-exchange code on a single GPU
    static const bool force_peer_instance = std::getenv("QSB_FORCE_PEER_INSTANCE") != nullptr;
    const bool peer = a.peer_mode || force_peer_instance, stream = a.n_in != 0ull || a.census_aos != nullptr;
    if (peer && stream) track_warpq_kernel<QSB_VALIDATION, 1, 1><<<grid, kWqThreads, sizeof(WqShared), s>>>(a);
    else if (peer)      track_warpq_kernel<QSB_VALIDATION, 1, 0><<<grid, kWqThreads, sizeof(WqShared), s>>>(a);
    else if (stream)    track_warpq_kernel<QSB_VALIDATION, 0, 1><<<grid, kWqThreads, sizeof(WqShared), s>>>(a);
    else                track_warpq_kernel<QSB_VALIDATION, 0, 0><<<grid, kWqThreads, sizeof(WqShared), s>>>(a);
}

// registers per thread, resident blocks per SM, threads per block, shared-memory bytes per block, particle slots per block
void QSB_EVT_ATTR_NAME(int* regs, int* max_blocks_per_sm, int* threads, int* smem_bytes, int* slots)
{
    const void* kernels[4] = { (const void*)track_warpq_kernel<QSB_VALIDATION, 0, 0>, (const void*)track_warpq_kernel<QSB_VALIDATION, 0, 1>,
                               (const void*)track_warpq_kernel<QSB_VALIDATION, 1, 0>, (const void*)track_warpq_kernel<QSB_VALIDATION, 1, 1> };
    int r = 0, blocks = 1 << 30;
    for (const void* k : kernels)
    {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WqShared));
        cudaFuncAttributes attr;
        if (cudaFuncGetAttributes(&attr, k) == cudaSuccess && attr.numRegs > r) r = attr.numRegs;
        int n = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, kWqThreads, sizeof(WqShared));
        if (n < blocks) blocks = n;
    }
    if (regs) *regs = r;
    if (max_blocks_per_sm) *max_blocks_per_sm = blocks;
    if (threads) *threads = kWqThreads;
    if (smem_bytes) *smem_bytes = (int)sizeof(WqShared);
    if (slots) *slots = kWqSlots;
}

} // namespace qsb
