// track_physics.cuh -- the physics and plumbing of one particle history on the device, shared by the two tracking kernels:
// the history-based persistent kernel (track_kernels.cu: one lane = one history held in registers) and the event-based one
// (track_event_kernels.cu: particles live in shared memory, warps run whole batches of one event type).  Everything here is
// __forceinline__ / file-local (anonymous namespace), so both translation units get their own copy, compiled with their own
// build flags (validation: --fmad=false + strict math; fast).  Reference map: see track_kernels.cu.
#ifndef QSB_TRACK_PHYSICS_CUH
#define QSB_TRACK_PHYSICS_CUH
#include <cuda_runtime.h>
#include <cstdint>

#include "device_types.cuh"
#include "../qs_rng.h"
#include "../qs_strict_math.h"

#ifndef QSB_VALIDATION
#define QSB_VALIDATION 1
#endif

namespace qsb {
namespace {

constexpr double kNeutronRestMassEnergy = 9.395656981095e+2;
constexpr double kSpeedOfLight = 2.99792458e+10;
constexpr double kTinyDouble = 1.0e-13;
constexpr double kSmallDouble = 1.0e-10;
constexpr double kHugeDouble = 1.0e+75;
constexpr unsigned kFullMask = 0xffffffffu;
constexpr unsigned long long kNoTicket = ~0ull;
#ifndef QSB_REFILL
#define QSB_REFILL 8
#endif
constexpr unsigned kRefill = QSB_REFILL;  // idle lanes a warp lets gather before it runs its service phase
#ifndef QSB_OPT_PREFETCH_L1
#define QSB_OPT_PREFETCH_L1 1
#endif
#ifndef QSB_OPT_PREFETCH_L2
#define QSB_OPT_PREFETCH_L2 1
#endif
#ifndef QSB_OPT_PREFETCH_ADJ
#define QSB_OPT_PREFETCH_ADJ 0
#endif
#ifndef QSB_OPT_SPARE
#define QSB_OPT_SPARE 1
#endif
#ifndef QSB_COLLIDE
#define QSB_COLLIDE 20
#endif
constexpr unsigned kCollide = QSB_COLLIDE; // lanes with a pending collision a warp lets gather before it runs the collision pass
constexpr unsigned kTicketBatch = 32;     // tickets a warp reserves per atomicAdd on the queue head

// facet -> 3 of the cell's 14 points, and facet -> matching facet of the face neighbour (src/MC_Domain.cc:41-50)
__constant__ int8_t c_facet_points[24][4] = {
    {1, 3, 8, 0},  {3, 7, 8, 0},  {7, 5, 8, 0},  {5, 1, 8, 0},  {0, 4, 9, 0},  {4, 6, 9, 0},  {6, 2, 9, 0},  {2, 0, 9, 0},
    {3, 2, 10, 0}, {2, 6, 10, 0}, {6, 7, 10, 0}, {7, 3, 10, 0}, {0, 1, 11, 0}, {1, 5, 11, 0}, {5, 4, 11, 0}, {4, 0, 11, 0},
    {4, 5, 12, 0}, {5, 7, 12, 0}, {7, 6, 12, 0}, {6, 4, 12, 0}, {0, 2, 13, 0}, {2, 3, 13, 0}, {3, 1, 13, 0}, {1, 0, 13, 0} };
// the facet of face `f` whose base is the face-rectangle edge e: 0 = low u, 1 = high u, 2 = low v, 3 = high v, where
// (u, v) are the two in-face axes in x<y<z order (derived from c_facet_points; checked by tests/test_host_model.py)
__constant__ int8_t c_facet_of_edge[6][4] = { {3, 1, 0, 2}, {4, 6, 7, 5}, {9, 11, 8, 10}, {15, 13, 12, 14}, {19, 17, 16, 18}, {20, 22, 23, 21} };

struct Particle
{
    double x, y, z, vx, vy, vz, alpha, beta, gamma;
    double energy, weight, ttc, age, nmfp, nseg, total_xs, speed;
    uint64_t seed, id;
    uint4 head;                 // first 16 bytes of the current cell's CellRec
    int cell, facet, group;
    int last_event, num_collisions, breed, species;
};

// Velocity of an in-flight particle.  Validation: the stored vector, exactly as the reference carries it.  Fast build:
// speed * direction cosine (equal up to rounding), so the three velocity registers are not live across the tracking loop.
#if QSB_VALIDATION
#define QSB_VX(p_) ((p_).vx)
#define QSB_VY(p_) ((p_).vy)
#define QSB_VZ(p_) ((p_).vz)
#else
#define QSB_VX(p_) ((p_).speed * (p_).alpha)
#define QSB_VY(p_) ((p_).speed * (p_).beta)
#define QSB_VZ(p_) ((p_).speed * (p_).gamma)
#endif

__device__ __forceinline__ int cell_ix(const uint4& h) { return (int)(h.x & 0xffffu); }
__device__ __forceinline__ int cell_iy(const uint4& h) { return (int)(h.x >> 16); }
__device__ __forceinline__ int cell_iz(const uint4& h) { return (int)(h.y & 0xffffu); }
__device__ __forceinline__ int cell_material(const uint4& h) { return (int)((h.y >> 16) & 0xffu); }
__device__ __forceinline__ int face_event(const uint4& h, int face) { return (int)((h.z >> (4 * face)) & 0xfu); }

// last_event tag of a raw fission secondary in the vault (see push_raw_child)
constexpr int kRawChild = 0x52415743;
// what a lane's in-flight particle needs next
enum { kStateIdle = 0, kStateSegment = 1, kStateCollision = 2, kStateTail = 3 };

struct Counters     // per-thread balance tallies, flushed once per kernel (src/Tallies.hh:36-100)
{
    unsigned int segments, collisions, absorbs, fissions, produced, escapes, census;
#if QSB_VALIDATION
    unsigned int scatters, lookups, slow, mismatch;     // fast build: scatters = collisions - absorbs - fissions; no diagnostics
#endif
};

// one facet plane {A,B,C,D}: two 16-byte read-only loads
__device__ __forceinline__ double4 load_plane(const double4* __restrict__ p)
{
    const double2 lo = __ldg(reinterpret_cast<const double2*>(p));
    const double2 hi = __ldg(reinterpret_cast<const double2*>(p) + 1);
    return make_double4(lo.x, lo.y, hi.x, hi.y);
}

// First 16 bytes of a cell record, read on entering the cell.  The record's second sector (facet codes, needed a hundred
// instructions into the next segment and only then addressable) is pulled into L1 alongside, so that load hits.
__device__ __forceinline__ uint4 load_cell_head(const DevImage& im, int cell)
{
    const char* rec = reinterpret_cast<const char*>(im.cells + cell);
#if QSB_OPT_PREFETCH_L1
    asm volatile("prefetch.global.L1 [%0];" :: "l"(rec + 32));
#endif
    return __ldg(reinterpret_cast<const uint4*>(rec));
}

// a freshly reserved ticket batch: start moving its particle records towards L2 (lane i takes ticket first + i)
__device__ __forceinline__ void prefetch_tickets(const TrackArgs& a, unsigned long long first, unsigned lane)
{
    const unsigned long long t = first + lane;
    if (t < a.n_in)
    {
        const char* rec = reinterpret_cast<const char*>(a.in_aos + t);
        asm volatile("prefetch.global.L2 [%0];" :: "l"(rec));
        asm volatile("prefetch.global.L2 [%0];" :: "l"(rec + 128));
    }
    else if (t - a.n_in < a.proc.capacity)
    {
        const unsigned long long i = t - a.n_in;
        const VaultView& v = a.proc;
        // one lane in four touches each 32-byte sector of the 8-byte arrays
        if ((lane & 3u) == 0u)
        {
            const double* f64[15] = { v.x, v.y, v.z, v.vx, v.vy, v.vz, v.energy, v.weight, v.ttc, v.age, v.nmfp, v.nseg, v.dirx, v.diry, v.dirz };
#pragma unroll
            for (int k = 0; k < 15; ++k) asm volatile("prefetch.global.L2 [%0];" :: "l"(f64[k] + i));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(v.seed + i));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(v.id + i));
        }
        if ((lane & 1u) == 0u) asm volatile("prefetch.global.L2 [%0];" :: "l"(v.tags + i));
        if ((lane & 7u) == 0u) asm volatile("prefetch.global.L2 [%0];" :: "l"(v.cell + i));
    }
}

// log / sin / cos: the portable functions of qs_strict_math.h in BOTH builds.  Validation (--fmad=false) gets the bits
// of the CPU oracle; the fast build contracts them to FMAs (same ~1 ulp accuracy) and avoids the CUDA math library's
// out-of-line argument-reduction slow path, which the tracking loop can never reach (0 <= phi < 2 pi, 0 < r < 1).
__device__ __forceinline__ double m_log(double x) { return qs_strict_log(x); }
__device__ __forceinline__ void m_sincos(double phi, double* s, double* c) { qs_strict_sincos(phi, s, c); }

// Arithmetic that differs between the two builds.  Validation: IEEE division and square root exactly as the reference
// (and the oracle) evaluate them.  Fast: the hardware reciprocal / reciprocal-square-root approximation refined by two
// Newton / Goldschmidt steps in FMA arithmetic (relative error ~1e-15, branch-free, a third of the instructions).
__device__ __forceinline__ double approx_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = __fma_rn(-x, r, 1.0);
    r = __fma_rn(r, e, r);
    e = __fma_rn(-x, r, 1.0);
    return __fma_rn(r, e, r);
}
__device__ __forceinline__ double approx_sqrt(double x)
{
    x = fmax(x, 1e-300);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = __fma_rn(-g, h, 0.5);
    g = __fma_rn(g, r, g); h = __fma_rn(h, r, h);
    r = __fma_rn(-g, h, 0.5);
    return __fma_rn(g, r, g);
}
#if QSB_VALIDATION
__device__ __forceinline__ double m_div(double a, double b) { return a / b; }
__device__ __forceinline__ double m_sqrt(double x) { return sqrt(x); }
#else
__device__ __forceinline__ double m_div(double a, double b) { return a * approx_rcp(b); }
__device__ __forceinline__ double m_sqrt(double x) { return approx_sqrt(x); }
#endif

// src/NuclearData.cc:208-227.  The reference bisects the nGroups+1 edges: for e[0] < energy <= e[n-1] it returns the
// largest i <= n-2 with e[i] <= energy.  The edges are log-spaced (src/NuclearData.cc:105-119), so the index is first
// estimated from a single-precision log2 and then corrected against the table itself -- the answer is decided by
// the same comparisons on the same doubles, so it is the reference's for ANY increasing table; a bad estimate only costs
// extra steps.  Two dependent loads instead of eight.
// Above the last edge the reference returns nGroups, one past the last group, and then indexes its tables with it
// (undefined behaviour on the host; here it would be a stray flux tally or an illegal address).  Reachable only when a
// deck's eMax is below the 20 MeV a fission neutron can carry (src/NuclearData.cc:77): the particle is put in the last
// group and counted, and qsb_track reports the cycle as failed.
__device__ __forceinline__ int energy_group(const TrackArgs& a, double energy)
{
    const DevImage& im = a.im;
    const int n = im.n_groups + 1;
    const double* __restrict__ e = im.energies;
    if (energy <= __ldg(e)) return 0;
    if (__builtin_expect(energy > __ldg(e + n - 1), 0)) { atomicAdd(&a.ctl->bad_group, 1u); return n - 2; }
    int i = (int)((__log2f((float)energy) - im.group_log2_lo) * im.group_inv_dlog2);
    i = min(max(i, 0), n - 2);
    while (i > 0 && energy < __ldg(e + i)) --i;
    while (i < n - 2 && energy >= __ldg(e + i + 1)) ++i;
    return i;
}

__device__ __forceinline__ double speed_of(const Particle& p) { return m_sqrt(p.vx * p.vx + p.vy * p.vy + p.vz * p.vz); }

// MC_Load_Particle + MC_Particle(const MC_Base_Particle&): src/MC_Load_Particle.cc:11-29,
// src/MC_Base_Particle.hh:287-331
__device__ __forceinline__ void reload_transform(const TrackArgs& a, Particle& p, double dt, bool derive_direction)
{
    p.speed = speed_of(p);
    if (derive_direction)
    {
        const double factor = m_div(1.0, p.speed);
        p.alpha = factor * p.vx; p.beta = factor * p.vy; p.gamma = factor * p.vz;
    }
    if (p.ttc <= 0.0) p.ttc += dt;
    if (p.age < 0.0) p.age = 0.0;
    p.group = energy_group(a, p.energy);
}

__device__ __forceinline__ int load_particle(const TrackArgs& a, unsigned long long i, Particle& p)
{
    const VaultView& v = a.proc;
    p.x = __ldcg(v.x + i); p.y = __ldcg(v.y + i); p.z = __ldcg(v.z + i);
    p.vx = __ldcg(v.vx + i); p.vy = __ldcg(v.vy + i); p.vz = __ldcg(v.vz + i);
    p.energy = __ldcg(v.energy + i); p.weight = __ldcg(v.weight + i); p.ttc = __ldcg(v.ttc + i);
    p.age = __ldcg(v.age + i); p.nmfp = __ldcg(v.nmfp + i); p.nseg = __ldcg(v.nseg + i);
    p.seed = (uint64_t)__ldcg(v.seed + i); p.id = (uint64_t)__ldcg(v.id + i);
    p.cell = __ldcg(v.cell + i);
    const int4 t = __ldcg(v.tags + i);
    p.last_event = t.x; p.num_collisions = t.y; p.breed = t.z; p.species = t.w;
    p.alpha = __ldcg(v.dirx + i); p.beta = __ldcg(v.diry + i); p.gamma = __ldcg(v.dirz + i);
    p.facet = 0; p.total_xs = 0.0;
    p.head = load_cell_head(a.im, p.cell);
    if (p.last_event == kRawChild) { p.last_event = QSB_EV_COLLISION; p.speed = 0.0; p.group = 0; return kStateTail; }
    reload_transform(a, p, a.dt, p.alpha != p.alpha);
    return kStateSegment;
}

// host-buffer streaming: ticket i is record i of the host vault, DMA-copied into HBM as it is (136-byte
// MC_Base_Particle layout, src/MC_Base_Particle.hh:75-92).  Read once per history with 17 L2 (.cg) loads.
__device__ __forceinline__ void load_particle_aos(const TrackArgs& a, unsigned long long i, Particle& p)
{
    const double* __restrict__ r = reinterpret_cast<const double*>(a.in_aos + i);
    p.x = __ldcg(r + 0); p.y = __ldcg(r + 1); p.z = __ldcg(r + 2);
    p.vx = __ldcg(r + 3); p.vy = __ldcg(r + 4); p.vz = __ldcg(r + 5);
    p.energy = __ldcg(r + 6); p.weight = __ldcg(r + 7); p.ttc = __ldcg(r + 8);
    p.age = __ldcg(r + 9); p.nmfp = __ldcg(r + 10); p.nseg = __ldcg(r + 11);
    const unsigned long long* __restrict__ u = reinterpret_cast<const unsigned long long*>(r);
    p.seed = (uint64_t)__ldcg(u + 12); p.id = (uint64_t)__ldcg(u + 13);
    const unsigned long long t0 = __ldcg(u + 14), t1 = __ldcg(u + 15), t2 = __ldcg(u + 16);
    p.last_event = (int)(unsigned)t0; p.num_collisions = (int)(unsigned)(t0 >> 32);
    p.breed = (int)(unsigned)t1; p.species = (int)(unsigned)(t1 >> 32);
    const int domain = (int)(unsigned)t2, cell = (int)(unsigned)(t2 >> 32);
    p.cell = __ldg(a.im.domain_cell_offset + domain) + cell;
    p.facet = 0; p.total_xs = 0.0;
    p.head = load_cell_head(a.im, p.cell);
    reload_transform(a, p, a.dt, true);
}

__device__ __forceinline__ void store_particle(const VaultView& v, unsigned long long i, const Particle& p, bool with_direction)
{
    __stcg(v.x + i, p.x); __stcg(v.y + i, p.y); __stcg(v.z + i, p.z);
    __stcg(v.vx + i, QSB_VX(p)); __stcg(v.vy + i, QSB_VY(p)); __stcg(v.vz + i, QSB_VZ(p));
    __stcg(v.energy + i, p.energy); __stcg(v.weight + i, p.weight); __stcg(v.ttc + i, p.ttc);
    __stcg(v.age + i, p.age); __stcg(v.nmfp + i, p.nmfp); __stcg(v.nseg + i, p.nseg);
    __stcg(v.seed + i, (unsigned long long)p.seed); __stcg(v.id + i, (unsigned long long)p.id);
    __stcg(v.cell + i, p.cell);
    __stcg(v.tags + i, make_int4(p.last_event, p.num_collisions, p.breed, p.species));
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    __stcg(v.dirx + i, with_direction ? p.alpha : nan);
    __stcg(v.diry + i, with_direction ? p.beta : nan);
    __stcg(v.dirz + i, with_direction ? p.gamma : nan);
}

// A deposit into a PEER's vault (NVLink stores, no ordering between them and no fence): the record carries the XOR of
// its own words and a per-launch salt in `check`, so the receiver can tell a complete record from one whose stores are
// still landing (or from the slot's previous contents) without the sender ever waiting for its stores to be acknowledged.
__device__ __forceinline__ unsigned long long bits(double v) { return (unsigned long long)__double_as_longlong(v); }
__device__ __forceinline__ void store_deposit(const VaultView& v, unsigned long long i, const Particle& p, int cell, uint32_t vault_epoch)
{
    const double vx = QSB_VX(p), vy = QSB_VY(p), vz = QSB_VZ(p);
    const unsigned long long t0 = (unsigned long long)(unsigned)p.last_event | ((unsigned long long)(unsigned)p.num_collisions << 32);
    const unsigned long long t1 = (unsigned long long)(unsigned)p.breed | ((unsigned long long)(unsigned)p.species << 32);
    unsigned long long x = deposit_salt(vault_epoch) ^ (unsigned long long)(unsigned)cell ^ t0 ^ t1 ^ p.seed ^ p.id;
    x ^= bits(p.x) ^ bits(p.y) ^ bits(p.z) ^ bits(vx) ^ bits(vy) ^ bits(vz) ^ bits(p.energy) ^ bits(p.weight) ^ bits(p.ttc);
    x ^= bits(p.age) ^ bits(p.nmfp) ^ bits(p.nseg) ^ bits(p.alpha) ^ bits(p.beta) ^ bits(p.gamma);
    __stcg(v.x + i, p.x); __stcg(v.y + i, p.y); __stcg(v.z + i, p.z);
    __stcg(v.vx + i, vx); __stcg(v.vy + i, vy); __stcg(v.vz + i, vz);
    __stcg(v.energy + i, p.energy); __stcg(v.weight + i, p.weight); __stcg(v.ttc + i, p.ttc);
    __stcg(v.age + i, p.age); __stcg(v.nmfp + i, p.nmfp); __stcg(v.nseg + i, p.nseg);
    __stcg(v.seed + i, (unsigned long long)p.seed); __stcg(v.id + i, (unsigned long long)p.id);
    __stcg(v.cell + i, cell);
    __stcg(v.tags + i, make_int4(p.last_event, p.num_collisions, p.breed, p.species));
    __stcg(v.dirx + i, p.alpha); __stcg(v.diry + i, p.beta); __stcg(v.dirz + i, p.gamma);
    __stcg(v.check + i, x);
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(v.ready + i), "r"(vault_epoch | kArrivalBit) : "memory");
}

// the receiving side: true when the record in slot i is complete
__device__ __forceinline__ bool deposit_complete(const VaultView& v, unsigned long long i, uint32_t vault_epoch)
{
    unsigned long long x = deposit_salt(vault_epoch) ^ (unsigned long long)(unsigned)__ldcg(v.cell + i);
    const int4 t = __ldcg(v.tags + i);
    x ^= (unsigned long long)(unsigned)t.x | ((unsigned long long)(unsigned)t.y << 32);
    x ^= (unsigned long long)(unsigned)t.z | ((unsigned long long)(unsigned)t.w << 32);
    x ^= __ldcg(v.seed + i) ^ __ldcg(v.id + i);
    const double* f64[15] = { v.x, v.y, v.z, v.vx, v.vy, v.vz, v.energy, v.weight, v.ttc, v.age, v.nmfp, v.nseg, v.dirx, v.diry, v.dirz };
#pragma unroll
    for (int k = 0; k < 15; ++k) x ^= bits(__ldcg(f64[k] + i));
    return x == __ldcg(v.check + i);
}

// ---- nearest facet, full path --------------------------------------------------------------------------

// ray / triangle test of one facet: src/MCT.cc:280-395
__device__ __forceinline__ double distance_to_segment(double plane_tolerance, double dot, const double4 pl,
                                                      const double* __restrict__ n0, const double* __restrict__ n1,
                                                      const double* __restrict__ n2, double px, double py, double pz,
                                                      double alpha, double beta, double gamma)
{
    const double bb_tol = 1e-9;
    const double numerator = -1.0 * (pl.x * px + pl.y * py + pl.z * pz + pl.w);
    if (numerator < 0.0 && numerator * numerator > plane_tolerance) return kHugeDouble;

    const double distance = numerator / dot;
    const double ix = px + distance * alpha;
    const double iy = py + distance * beta;
    const double iz = pz + distance * gamma;

    const double ax = __ldg(n0), ay = __ldg(n0 + 1), az = __ldg(n0 + 2);
    const double bx = __ldg(n1), by = __ldg(n1 + 1), bz = __ldg(n1 + 2);
    const double cx = __ldg(n2), cy = __ldg(n2 + 1), cz = __ldg(n2 + 2);

#define QSB_BELOW(a_, b_, c_, i_) ((a_) > (i_) + bb_tol && (b_) > (i_) + bb_tol && (c_) > (i_) + bb_tol)
#define QSB_ABOVE(a_, b_, c_, i_) ((a_) < (i_) - bb_tol && (b_) < (i_) - bb_tol && (c_) < (i_) - bb_tol)
#define QSB_CROSS(ax_, ay_, bx_, by_, cx_, cy_) (((bx_) - (ax_)) * ((cy_) - (ay_)) - ((by_) - (ay_)) * ((cx_) - (ax_)))

    double cross0 = 0, cross1 = 0, cross2 = 0;
    if (pl.z < -0.5 || pl.z > 0.5)
    {
        if (QSB_BELOW(ax, bx, cx, ix) || QSB_ABOVE(ax, bx, cx, ix) || QSB_BELOW(ay, by, cy, iy) || QSB_ABOVE(ay, by, cy, iy))
            return kHugeDouble;
        cross1 = QSB_CROSS(ax, ay, bx, by, ix, iy);
        cross2 = QSB_CROSS(bx, by, cx, cy, ix, iy);
        cross0 = QSB_CROSS(cx, cy, ax, ay, ix, iy);
    }
    else if (pl.y < -0.5 || pl.y > 0.5)
    {
        if (QSB_BELOW(ax, bx, cx, ix) || QSB_ABOVE(ax, bx, cx, ix) || QSB_BELOW(az, bz, cz, iz) || QSB_ABOVE(az, bz, cz, iz))
            return kHugeDouble;
        cross1 = QSB_CROSS(az, ax, bz, bx, iz, ix);
        cross2 = QSB_CROSS(bz, bx, cz, cx, iz, ix);
        cross0 = QSB_CROSS(cz, cx, az, ax, iz, ix);
    }
    else if (pl.x < -0.5 || pl.x > 0.5)
    {
        if (QSB_BELOW(az, bz, cz, iz) || QSB_ABOVE(az, bz, cz, iz) || QSB_BELOW(ay, by, cy, iy) || QSB_ABOVE(ay, by, cy, iy))
            return kHugeDouble;
        cross1 = QSB_CROSS(ay, az, by, bz, iy, iz);
        cross2 = QSB_CROSS(by, bz, cy, cz, iy, iz);
        cross0 = QSB_CROSS(cy, cz, ay, az, iy, iz);
    }
#undef QSB_BELOW
#undef QSB_ABOVE
#undef QSB_CROSS

    const double cross_tol = 1e-9 * fabs(cross0 + cross1 + cross2);
    if ((cross0 > -cross_tol && cross1 > -cross_tol && cross2 > -cross_tol) ||
        (cross0 <  cross_tol && cross1 <  cross_tol && cross2 <  cross_tol))
        return distance;
    return kHugeDouble;
}

// all 24 facets of the cell, nearest positive hit, fallback + retry nudge: src/MCT.cc:436-621, :87-137.
// Cold path (and the only path when the mesh is not compact-encodable); may move the coordinate.
__device__ __noinline__ void nearest_facet_full(const double4* __restrict__ planes, const double* __restrict__ nodes,
                                                double* px, double* py, double* pz, double alpha, double beta, double gamma,
                                                double nseg, int* out_facet, double* out_distance)
{
    int iteration = 0;
    double move_factor = 0.5 * kSmallDouble;
    int nf_facet; double nf_distance;
    double x = *px, y = *py, z = *pz;
    for (;;)
    {
        const double plane_tolerance = 1e-16 * (x * x + y * y + z * z);
        nf_facet = 0; nf_distance = 1e80;
        int neg_facet = 0; double neg_distance = -kHugeDouble;
#pragma unroll 1
        for (int f = 0; f < 24; ++f)
        {
            double t = kHugeDouble;
            const double4 pl = load_plane(planes + f);
            const double dot = (pl.x * alpha + pl.y * beta + pl.z * gamma);
            if (dot > 0.0)
                t = distance_to_segment(plane_tolerance, dot, pl, nodes + 3 * c_facet_points[f][0],
                                        nodes + 3 * c_facet_points[f][1], nodes + 3 * c_facet_points[f][2], x, y, z, alpha, beta, gamma);
            // MCT_Nearest_Facet_Find_Nearest folded into the loop: same order, same comparisons
            if (t > 0.0) { if (t <= nf_distance) { nf_distance = t; nf_facet = f; } }
            else if (t > neg_distance) { neg_distance = t; neg_facet = f; }
        }
        if (nf_distance == kHugeDouble && neg_distance != -kHugeDouble) { nf_distance = neg_distance; nf_facet = neg_facet; }

        bool retry = false;
        if ((nf_distance == kHugeDouble && move_factor > 0) || (nseg > 10000000 && nf_distance <= 0.0))
        {
            double mx = 0, my = 0, mz = 0;
            for (int k = 0; k < 14; ++k) { mx += __ldg(nodes + 3 * k); my += __ldg(nodes + 3 * k + 1); mz += __ldg(nodes + 3 * k + 2); }
            const double inv = 1.0 / ((double)14);
            mx *= inv; my *= inv; mz *= inv;
            x += move_factor * (mx - x);
            y += move_factor * (my - y);
            z += move_factor * (mz - z);
            iteration++;
            move_factor *= 2.0;
            if (move_factor > 1.0e-2) move_factor = 1.0e-2;
            retry = iteration != 10000;
        }
        if (!retry) break;
    }
    if (nf_distance < 0) nf_distance = 0;
    *px = x; *py = y; *pz = z;
    *out_facet = nf_facet; *out_distance = nf_distance;
}

#if !QSB_VALIDATION
// ---- nearest facet, fast build ---------------------------------------------------------------------------
// The cell is an axis-aligned box and the fast build only has to be statistically equivalent to the reference (its
// arithmetic already differs in the last bits), so the exit is the face with the smallest gap / |direction| and the
// distance is that quotient -- no facet code, no plane, no triangle on the face (reflection and adjacency only need the
// face).  A particle that rounding has left a hair outside its cell sees a negative gap and crosses with a zero-length
// segment, which is what the reference's negative-distance fallback + clamp does (src/MCT.cc:468-476, :114).
__device__ __forceinline__ bool nearest_facet_fast(const DevImage& im, const Particle& p, int& facet, double& distance)
{
    const double x0 = cell_ix(p.head) * im.dx, y0 = cell_iy(p.head) * im.dy, z0 = cell_iz(p.head) * im.dz;
    const double gx = p.alpha > 0 ? (x0 + im.dx) - p.x : p.x - x0, ax = fabs(p.alpha);
    const double gy = p.beta  > 0 ? (y0 + im.dy) - p.y : p.y - y0, ay = fabs(p.beta);
    const double gz = p.gamma > 0 ? (z0 + im.dz) - p.z : p.z - z0, az = fabs(p.gamma);
    int w = -1; double gw = 0, aw = 1;
    if (ax > 0) { w = 0; gw = gx; aw = ax; }
    if (ay > 0 && (w < 0 || gy * aw < gw * ay)) { w = 1; gw = gy; aw = ay; }
    if (az > 0 && (w < 0 || gz * aw < gw * az)) { w = 2; gw = gz; aw = az; }
    if (w < 0) return false;
    const double dw = w == 0 ? p.alpha : (w == 1 ? p.beta : p.gamma);
    const int face = 2 * w + (dw > 0 ? 0 : 1);
    facet = 4 * face;
#if QSB_OPT_PREFETCH_ADJ
    // every other segment ends on this face: start pulling the neighbour's record towards L1 now (the adjacency word sits in
    // the sectors of this cell's record that are already there), instead of a dependent miss after the crossing
    asm volatile("prefetch.global.L1 [%0];" :: "l"(im.cells + __ldg(im.cells[p.cell].adj + face)));
#endif
    distance = fmax(gw * approx_rcp(aw), 0.0);
    return true;
}
#else
// ---- nearest facet, filtered fast path -----------------------------------------------------------------
// Returns false when the configuration is within the safety margin of anything the reference treats with
// tolerances (cell edges, face diagonals, the exit face itself, a particle outside its cell); the caller then
// takes the full path.  When it returns true, (facet, distance) carry exactly the bits of the full path.
__device__ __forceinline__ bool nearest_facet_fast(const DevImage& im, const Particle& p, int& facet, double& distance)
{
    const int ix = cell_ix(p.head), iy = cell_iy(p.head), iz = cell_iz(p.head);
    // exact node coordinates of the cell's corners: index * cell size (src/GlobalFccGrid.cc:112-131)
    const double x0 = ix * im.dx, x1 = (ix + 1) * im.dx;
    const double y0 = iy * im.dy, y1 = (iy + 1) * im.dy;
    const double z0 = iz * im.dz, z1 = (iz + 1) * im.dz;
    const double m = im.margin;
    bool ok = p.x >= x0 - m && p.x <= x1 + m && p.y >= y0 - m && p.y <= y1 + m && p.z >= z0 - m && p.z <= z1 + m;

    // gap to the candidate face of each axis (the face the direction points at) and |direction|
    const double gx = p.alpha > 0 ? x1 - p.x : p.x - x0, ax = fabs(p.alpha);
    const double gy = p.beta  > 0 ? y1 - p.y : p.y - y0, ay = fabs(p.beta);
    const double gz = p.gamma > 0 ? z1 - p.z : p.z - z0, az = fabs(p.gamma);
    // exit axis = argmin gap/|dir| over axes with dir != 0, by cross multiplication
    int w = -1; double gw = 0, aw = 1;
    if (ax > 0) { w = 0; gw = gx; aw = ax; }
    if (ay > 0 && (w < 0 || gy * aw < gw * ay)) { w = 1; gw = gy; aw = ay; }
    if (az > 0 && (w < 0 || gz * aw < gw * az)) { w = 2; gw = gz; aw = az; }
    if (w < 0) return false;
    ok = ok && gw > m;
    const double t = gw * approx_rcp(aw);          // approximate distance (~1e-15 relative): only used for the filter
    const double ex = p.x + t * p.alpha, ey = p.y + t * p.beta, ez = p.z + t * p.gamma;

    // normalised in-face coordinates of the exit point, (u, v) = the two other axes in x<y<z order
    double su, sv, pw, dw, c_lo, c_hi;
    if (w == 0)      { su = (ey - (y0 + 0.5 * im.dy)) * im.inv_hy; sv = (ez - (z0 + 0.5 * im.dz)) * im.inv_hz; pw = p.x; dw = p.alpha; c_lo = x0; c_hi = x1; }
    else if (w == 1) { su = (ex - (x0 + 0.5 * im.dx)) * im.inv_hx; sv = (ez - (z0 + 0.5 * im.dz)) * im.inv_hz; pw = p.y; dw = p.beta;  c_lo = y0; c_hi = y1; }
    else             { su = (ex - (x0 + 0.5 * im.dx)) * im.inv_hx; sv = (ey - (y0 + 0.5 * im.dy)) * im.inv_hy; pw = p.z; dw = p.gamma; c_lo = z0; c_hi = z1; }
    const double au = fabs(su), av = fabs(sv);
    const double mm = 1e-6;
    ok = ok && fmax(au, av) < 1.0 - mm && fabs(au - av) > mm;
    if (!ok) return false;

    const int face = 2 * w + (dw > 0 ? 0 : 1);
    const int edge = au > av ? (su > 0 ? 1 : 0) : (sv > 0 ? 3 : 2);
    const int f = c_facet_of_edge[face][edge];

    // rebuild the facet's exact plane from its code and evaluate the reference's expression for this facet only:
    // numerator = -1.0 * (A*x + B*y + C*z + D) and dot = A*alpha + B*beta + C*gamma with B = C = (+-)0 reduce to
    // the non-zero axis term (adding a signed zero to a non-zero double is exact)
    const unsigned code = __ldg(im.cells[p.cell].code + f);
    const double sign = (face & 1) ? -1.0 : 1.0;
    const double normal = sign * ((code & 1u) ? __longlong_as_double(0x3FEFFFFFFFFFFFFFll) : 1.0);
    const double coord = (face & 1) ? c_lo : c_hi;
    const int k = (int)((code >> 1) & 3u);
    const double d_abs = __longlong_as_double(__double_as_longlong(coord) + (k == 1 ? 1ll : (k == 2 ? -1ll : 0ll)));
    const double D = (face & 1) ? d_abs : -d_abs;
    const double numerator = -1.0 * (normal * pw + D);
    const double dot = normal * dw;
    const double dist = m_div(numerator, dot);
    if (!(fabs(dist - t) <= 1e-9 * (t + m))) return false;   // also catches NaN
    facet = f; distance = dist;
    return true;
}

#endif

// ---- segment outcome: 0 collision, 1 facet crossing, 2 census -----------------------------------------
__device__ __forceinline__ int segment_outcome(const TrackArgs& a, Particle& p, Counters& c)
{
    const DevImage& im = a.im;
    const double particle_speed = p.speed;

    bool force_collision = false;
    if (p.nmfp < 0.0) { force_collision = true; p.nmfp = kSmallDouble; }

    // weightedMacroscopicCrossSection (src/MacroscopicCrossSection.cc:59-80): the reference's per-cell cache holds
    // the same number for every cell of a material; {total, 1/total} are precomputed per (material, group)
    const double2 xs = __ldg(im.xs_pair + (size_t)cell_material(p.head) * im.n_groups + p.group);
    p.total_xs = xs.x;
    const double mean_free_path = (xs.x == 0.0) ? kHugeDouble : xs.y;

    if (p.nmfp == 0.0)
    {
        const double r = qs_rng_sample(&p.seed);
        p.nmfp = -1.0 * m_log(r);
    }

    double d_collision = force_collision ? kSmallDouble : p.nmfp * mean_free_path;
    double d_census = particle_speed * p.ttc;

    int nf_facet = 0; double d_facet = 0.0;
    const bool fast = im.compact && nearest_facet_fast(im, p, nf_facet, d_facet);
#if QSB_VALIDATION
    if (!fast || (a.check_mode & 1))
    {
        int f2; double d2;
        double qx = p.x, qy = p.y, qz = p.z;
        nearest_facet_full(im.planes + (size_t)p.cell * 24, im.nodes + (size_t)p.cell * 42, &qx, &qy, &qz,
                           p.alpha, p.beta, p.gamma, p.nseg, &f2, &d2);
        if (fast) { if (f2 != nf_facet || d2 != d_facet || qx != p.x || qy != p.y || qz != p.z) c.mismatch++; }
        else { c.slow++; }
        p.x = qx; p.y = qy; p.z = qz;
        nf_facet = f2; d_facet = d2;
    }
#else
    if (!fast)          // a zero direction vector (the fast build is only ever launched on the uniform brick grid: qsb_create)
    {
        // The reference's 24-facet search finds no facet for it either and ends in its error path; here the particle simply has
        // no facet to cross, and the segment is counted so that qsb_track can report it.  (The full search is not part of the
        // fast build: called from here, with the particle live, it alone set the kernel's register floor.)
        nf_facet = 0; d_facet = kHugeDouble;
        atomicAdd(&a.ctl->slow_geometry, 1ull);
    }
#endif
    if (force_collision) { d_facet = kHugeDouble; d_census = kHugeDouble; d_collision = kTinyDouble; }

    // MC_Find_Min: strict <, ties to the lower index
    int outcome = 0; double dmin = d_collision;
    if (d_facet < dmin) { dmin = d_facet; outcome = 1; }
    if (d_census < dmin) { dmin = d_census; outcome = 2; }

    const double segment_path_length = dmin;
#if QSB_VALIDATION
    p.nmfp -= segment_path_length / mean_free_path;
#else
    p.nmfp -= segment_path_length * ((xs.x == 0.0) ? 1.0 / kHugeDouble : xs.x);
#endif
    p.last_event = outcome == 0 ? QSB_EV_COLLISION : (outcome == 1 ? QSB_EV_FACET_TRANSIT : QSB_EV_CENSUS);
    if (outcome == 0) p.nmfp = 0.0;
    else if (outcome == 1) p.facet = nf_facet;
    else p.ttc = (p.ttc < 0.0) ? p.ttc : 0.0;
    if (force_collision) p.nmfp = 0.0;

    if (segment_path_length == 0.0) return outcome;

    p.x += (p.alpha * segment_path_length);
    p.y += (p.beta * segment_path_length);
    p.z += (p.gamma * segment_path_length);
    const double segment_path_time = m_div(segment_path_length, particle_speed);
    p.ttc -= segment_path_time;
    p.age += segment_path_time;
    if (p.ttc < 0.0) p.ttc = 0.0;

    // scalar flux tally (src/Tallies.hh:351-354): fire-and-forget f64 reduction in L2
    atomicAdd(a.flux + (size_t)p.cell * im.n_groups + p.group, segment_path_length * p.weight);
    return outcome;
}

// ---- collision ------------------------------------------------------------------------------------------
// updateTrajectory (src/CollisionEvent.cc:25-45) + DirectionCosine::Rotate3DVector (src/DirectionCosine.hh:123-146).
// Returns the speed the new velocity was built from.
__device__ __forceinline__ double update_trajectory(double energy, double angle, Particle& p)
{
    p.energy = energy;
    const double cosTheta = angle;
    double r = qs_rng_sample(&p.seed);
    const double phi = 2 * 3.14159265 * r;
    double sinPhi, cosPhi;
    m_sincos(phi, &sinPhi, &cosPhi);
    const double sinTheta = m_sqrt((1.0 - (cosTheta * cosTheta)));

    const double cos_theta = p.gamma;
    const double sin_theta = m_sqrt((1.0 - (cos_theta * cos_theta)));
    double cos_phi, sin_phi;
    if (sin_theta < 1e-6) { cos_phi = 1.0; sin_phi = 0.0; }
    else
    {
#if QSB_VALIDATION
        cos_phi = p.alpha / sin_theta; sin_phi = p.beta / sin_theta;
#else
        const double inv = approx_rcp(sin_theta);
        cos_phi = p.alpha * inv; sin_phi = p.beta * inv;
#endif
    }
    const double na =  cos_theta * cos_phi * (sinTheta * cosPhi) - sin_phi * (sinTheta * sinPhi) + sin_theta * cos_phi * cosTheta;
    const double nb =  cos_theta * sin_phi * (sinTheta * cosPhi) + cos_phi * (sinTheta * sinPhi) + sin_theta * sin_phi * cosTheta;
    const double ng = -sin_theta           * (sinTheta * cosPhi) +                                 cos_theta           * cosTheta;
    p.alpha = na; p.beta = nb; p.gamma = ng;

#if QSB_VALIDATION
    const double speed = (kSpeedOfLight *
                          sqrt((1.0 - ((kNeutronRestMassEnergy * kNeutronRestMassEnergy) /
                                       ((energy + kNeutronRestMassEnergy) * (energy + kNeutronRestMassEnergy))))));
#else
    const double ratio = kNeutronRestMassEnergy * approx_rcp(energy + kNeutronRestMassEnergy);
    const double speed = kSpeedOfLight * approx_sqrt(1.0 - ratio * ratio);
#endif
    p.vx = speed * p.alpha; p.vy = speed * p.beta; p.vz = speed * p.gamma;
    r = qs_rng_sample(&p.seed);
    p.nmfp = -1.0 * m_log(r);
    return speed;
}

// A fission secondary is appended to the processing vault "raw": the parent's state at the collision, the child's own
// random-number stream, and the outgoing energy / scattering cosine sampled for it (in the energy and nmfp fields), tagged
// kRawChild.  The reference computes the child's updateTrajectory right here in the parent's thread
// (src/CollisionEvent.cc:125-133); that is two draws from the CHILD's stream and touches nothing of the parent, so it is
// done instead by whichever lane loads the record, inside the converged collision-tail pass -- same arithmetic, same
// bits, but not a 300-instruction detour with one lane active.  Slots are handed out per warp (one atomic for all the
// secondaries of a collision pass), records are written at once and published one pass later (publish_children).

__device__ __forceinline__ void write_raw_child(const TrackArgs& a, unsigned long long i, const Particle& parent, uint64_t child_seed,
                                                double energy_out, double angle_out)
{
    const VaultView& v = a.proc;
    __stcg(v.x + i, parent.x); __stcg(v.y + i, parent.y); __stcg(v.z + i, parent.z);
#if QSB_VALIDATION
    __stcg(v.vx + i, parent.vx); __stcg(v.vy + i, parent.vy); __stcg(v.vz + i, parent.vz);
#endif                          // fast build: the child's velocity is rebuilt by its collision tail before anything reads it
    __stcg(v.energy + i, energy_out); __stcg(v.weight + i, parent.weight); __stcg(v.ttc + i, parent.ttc);
    __stcg(v.age + i, parent.age); __stcg(v.nmfp + i, angle_out); __stcg(v.nseg + i, parent.nseg);
    __stcg(v.seed + i, (unsigned long long)child_seed); __stcg(v.id + i, (unsigned long long)child_seed);
    __stcg(v.cell + i, parent.cell);
    __stcg(v.tags + i, make_int4(kRawChild, parent.num_collisions, parent.breed, parent.species));
    __stcg(v.dirx + i, parent.alpha); __stcg(v.diry + i, parent.beta); __stcg(v.dirz + i, parent.gamma);
}

// Publish the secondaries this lane wrote in an earlier pass (slots [first, first + n)): one release store -- MEMBAR.ALL.GPU
// + STG, no L1 invalidation (__threadfence() would add CCTL.IVALL and throw away the SM's cached cell records and tables) --
// then relaxed stores for the rest.  Done one pass late and for the whole warp at once: by then the record stores have long
// been acknowledged, so the barrier no longer waits a DRAM round trip per fission (measured: 2.4 us each, 12 % of all stall
// samples, when every fissioning lane published on the spot).
__device__ __forceinline__ void publish_children(const TrackArgs& a, unsigned long long first, unsigned n, uint32_t epoch)
{
    if (n == 0u) return;
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(a.proc.ready + first), "r"(epoch) : "memory");
    for (unsigned k = 1; k < n; ++k)
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(a.proc.ready + first + k), "r"(epoch) : "memory");
}

// The reference walks the material's (isotope, reaction) table subtracting each macroscopic cross section until
// the running value goes negative (src/CollisionEvent.cc:59-83).  Every isotope of a material carries the same
// reaction table (src/initMC.cc:160-196; checked per material on the host), so the NR per-reaction values are
// held in registers.  Returns the flat index iso * n_react + react, or -1.
//
// exact chain: the reference's own sequence of subtractions, isotope after isotope.
#if QSB_VALIDATION
template <int NR>
__device__ __noinline__ int select_reaction_chain(const double* __restrict__ table, int n_iso, int n_react, double current)
{
    double v[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) v[k] = (k < n_react) ? __ldg(table + k) : 0.0;
    for (int iso = 0; iso < n_iso; ++iso)
    {
        double run[NR];
        double cur = current;
#pragma unroll
        for (int k = 0; k < NR; ++k) { cur = cur - v[k]; run[k] = cur; }
        if (cur < 0)
        {
            // values are non-negative, so the chain is non-increasing: the first negative entry is the selection
            int first = NR - 1;
#pragma unroll
            for (int k = NR - 2; k >= 0; --k) if (run[k] < 0) first = k;
            return iso * n_react + first;
        }
        current = cur;
    }
    return -1;
}

// filtered exact selection: the chain value after k subtractions differs from (current - prefix sum) by at most
// k roundings of at most ulp(total)/2 each, i.e. by less than 180 * 2^-53 * total ~ 2e-14 * total.  The entry the chain
// selects is therefore the first k with current < prefix[k+1] whenever `current` keeps a distance of `guard`
// = 1e-11 * total from the two prefix sums around it -- which is decided here with one division instead of
// walking up to n_iso * NR dependent subtractions.  Inside the guard band (probability ~1e-9 per collision) or when
// the estimate falls outside the table, the exact chain decides.  tracking_mode bit 2 runs both and counts
// disagreements (tests require 0).
template <int NR>
__device__ __forceinline__ int select_reaction_periodic(const double* __restrict__ table, int n_iso, int n_react, double current,
                                                        double total, bool check, unsigned int& mismatch)
{
    double prefix[NR];
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < NR; ++k) { sum += (k < n_react) ? __ldg(table + k) : 0.0; prefix[k] = sum; }
    const double guard = 1e-11 * total;
    int selected = -2;
    if (sum > 0.0)
    {
        const int iso = (int)(current / sum);
        const double rest = current - (double)iso * sum;          // position inside isotope `iso`, up to a few ulp(total)
        if (iso < n_iso && rest > guard)
        {
            int first = -1;
            double below = 0.0, above = 0.0;
#pragma unroll
            for (int k = 0; k < NR; ++k)
            {
                if (first < 0 && k < n_react)
                {
                    if (rest < prefix[k]) { first = k; above = prefix[k]; }
                    else below = prefix[k];
                }
            }
            // `first` is the first entry whose prefix sum exceeds rest (a zero-width entry, cross section 0, is skipped
            // exactly as the chain skips it); accepted when rest sits clear of both neighbouring prefix sums
            if (first >= 0 && above - rest > guard && rest - below > guard) selected = iso * n_react + first;
        }
    }
    if (selected == -2) return select_reaction_chain<NR>(table, n_iso, n_react, current);
    if (check && select_reaction_chain<NR>(table, n_iso, n_react, current) != selected) mismatch++;
    return selected;
}

#else
// Fast build.  With identical isotopes the position inside ONE isotope's table decides the reaction, and the total is
// n_iso times that table's sum up to rounding: the uniform number r picks isotope floor(r * n_iso) and the fraction
// picks the reaction -- no division, no chain.  Differs from the reference's subtraction chain only where r * total
// lies within rounding of a table boundary.
__device__ __forceinline__ int select_reaction_fastbuild(const double* __restrict__ table, int n_iso, int n_react, double r)
{
    double prefix[9];
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < 9; ++k) { sum += (k < n_react) ? __ldg(table + k) : 0.0; prefix[k] = sum; }
    const double t = r * (double)n_iso;
    int iso = (int)t;
    if (iso >= n_iso) iso = n_iso - 1;
    const double rest = (t - (double)iso) * sum;
    int first = n_react - 1;
#pragma unroll
    for (int k = 7; k >= 0; --k) if (k < n_react && rest < prefix[k]) first = k;
    return iso * n_react + first;
}
#endif

__device__ __forceinline__ int select_reaction_generic(const double* __restrict__ table, int n_total, double current)
{
    for (int k = 0; k < n_total; ++k)
    {
        current -= __ldg(table + k);
        if (current < 0) return k;
    }
    return -1;
}

// Collision, first half (src/CollisionEvent.cc:50-133): pick the reaction, sample its outcome, tally, spawn the
// outgoing particles.  Returns their number (0: the history ends here) and their (energy, scattering cosine) pairs: pair 0
// is the parent's own, for collision_tail; pairs 1.. belong to the secondaries the caller spawns.
__device__ __forceinline__ int collision_head(const TrackArgs& a, Particle& p, Counters& c, double& energy0, double& angle0,
                                               double& energy1, double& angle1, double& energy2, double& angle2, double& energy3, double& angle3)
{
    const DevImage& im = a.im;
    const int mat = cell_material(p.head);
    // the material's (isotope, reaction) row for this group; with the compact table only isotope 0's entries exist, which is
    // all the periodic selections read
    const double* __restrict__ table = im.xs_compact ? im.xs_compact + ((size_t)mat * im.n_groups + p.group) * im.compact_react
                                                     : im.xs_react + ((size_t)mat * im.n_groups + p.group) * im.max_react;
    const int n_iso = __ldg(im.mat_n_iso + mat), n_react = __ldg(im.mat_n_react + mat);

    double r = qs_rng_sample(&p.seed);
    const double current = p.total_xs * r;
    int selected;
    const bool periodic = __ldg(im.mat_periodic + mat) != 0;
#if QSB_VALIDATION
    const bool check = (a.check_mode & 2) != 0;
    if (periodic && n_react <= 3)      selected = select_reaction_periodic<3>(table, n_iso, n_react, current, p.total_xs, check, c.mismatch);
    else if (periodic && n_react <= 9) selected = select_reaction_periodic<9>(table, n_iso, n_react, current, p.total_xs, check, c.mismatch);
    else                               selected = select_reaction_generic(table, n_iso * n_react, current);
    c.lookups += (selected < 0 ? n_iso * n_react : selected + 1);
#else
    if (periodic && n_react <= 9) selected = select_reaction_fastbuild(table, n_iso, n_react, r);
    else                          selected = select_reaction_generic(table, n_iso * n_react, current);
#endif
    if (selected < 0) { atomicAdd(&a.ctl->bad_reaction, 1u); return 0; }

    // NuclearDataReaction::sampleCollision (src/NuclearData.cc:54-88)
    double energyOut[4] = { 0.0, 0.0, 0.0, 0.0 }, angleOut[4] = { 0.0, 0.0, 0.0, 0.0 };
    int nOut = 0;
    const int rtype = __ldg(im.mat_react_type + (size_t)mat * im.max_react + selected);
    if (rtype == QSB_REACT_SCATTER)
    {
        nOut = 1;
        r = qs_rng_sample(&p.seed);
        energyOut[0] = p.energy * (1.0 - (r * __ldg(im.mat_inv_mass + mat)));     // r * (1.0 / mass), the quotient formed on the host
        r = qs_rng_sample(&p.seed) * 2.0 - 1.0;
        angleOut[0] = r;
    }
    else if (rtype == QSB_REACT_FISSION)
    {
        int n = (int)(__ldg(im.mat_nu_bar + mat) + qs_rng_sample(&p.seed));
        if (n > 4) n = 4;
        nOut = n;
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            if (i < n)
            {
                r = qs_rng_sample(&p.seed) / 2.0 + 0.5;
                energyOut[i] = (20 * r * r);
                r = qs_rng_sample(&p.seed) * 2.0 - 1.0;
                angleOut[i] = r;
            }
        }
    }

    c.collisions++;
#if QSB_VALIDATION
    if (rtype == QSB_REACT_SCATTER) c.scatters++;
    else
#endif
    if (rtype == QSB_REACT_ABSORPTION) c.absorbs++;
    else if (rtype == QSB_REACT_FISSION) { c.fissions++; c.produced += nOut; }

    energy0 = energyOut[0]; angle0 = angleOut[0];
    energy1 = energyOut[1]; angle1 = angleOut[1];
    energy2 = energyOut[2]; angle2 = angleOut[2];
    energy3 = energyOut[3]; angle3 = angleOut[3];
    return nOut;
}

// Collision, second half: the outgoing particle's new trajectory (src/CollisionEvent.cc:135-145).  `requeued` marks a
// particle that the reference stores as a base particle and loads again before it flies on -- the fissioning parent
// (src/CollisionEvent.cc:137-142) and every secondary (raw child records) -- i.e. MC_Load_Particle's transform on top:
// direction cosine re-derived from the velocity, census clock / age fixed up (parity hazards 1-2 of SURVEY.md 8a).
__device__ __forceinline__ void collision_tail(const TrackArgs& a, Particle& p, double energy0, double angle0, bool requeued)
{
    const double speed = update_trajectory(energy0, angle0, p);
#if QSB_VALIDATION
    (void)speed;
    p.speed = speed_of(p);
    if (requeued)
    {
        const double factor = 1.0 / p.speed;
        p.alpha = factor * p.vx; p.beta = factor * p.vy; p.gamma = factor * p.vz;
    }
#else
    p.speed = speed;            // |velocity| up to rounding; the direction cosine is already a unit vector
#endif
    if (requeued)
    {
        if (p.ttc <= 0.0) p.ttc += a.dt;
        if (p.age < 0.0) p.age = 0.0;
    }
    p.group = energy_group(a, p.energy);
}

// ---- facet crossing --------------------------------------------------------------------------------------
__device__ __forceinline__ void reflect_particle(const DevImage& im, Particle& p)
{
#if !QSB_VALIDATION
    if (im.compact)
    {
        // axis-aligned face: dir -= 2 (dir . n) n flips the component along the face normal (src/MCT.cc:401-429); the
        // particle left through this face, so it is heading outwards and the reference's dot > 0 test holds
        const int axis = p.facet >> 3;
        if (axis == 0) p.alpha = -p.alpha; else if (axis == 1) p.beta = -p.beta; else p.gamma = -p.gamma;
        return;
    }
#endif
    const double4 pl = load_plane(im.planes + (size_t)p.cell * 24 + p.facet);
    const double dot = 2.0 * (p.alpha * pl.x + p.beta * pl.y + p.gamma * pl.z);
    if (dot > 0)
    {
        p.alpha -= dot * pl.x;
        p.beta  -= dot * pl.y;
        p.gamma -= dot * pl.z;
    }
    const double speed = p.speed;
    p.vx = speed * p.alpha; p.vy = speed * p.beta; p.vz = speed * p.gamma;
    p.speed = speed_of(p);
}

__device__ __forceinline__ int flat_to_domain(const DevImage& im, int flat)
{
    int d = 0;
    while (d + 1 < im.n_domains && flat >= __ldg(im.domain_cell_offset + d + 1)) d++;
    return d;
}

__device__ __forceinline__ void fill_base(const DevImage& im, const Particle& p, qsb_base_particle& b)
{
    b.coordinate[0] = p.x; b.coordinate[1] = p.y; b.coordinate[2] = p.z;
    b.velocity[0] = QSB_VX(p); b.velocity[1] = QSB_VY(p); b.velocity[2] = QSB_VZ(p);
    b.kinetic_energy = p.energy; b.weight = p.weight; b.time_to_census = p.ttc; b.age = p.age;
    b.num_mean_free_paths = p.nmfp; b.num_segments = p.nseg;
    b.random_number_seed = p.seed; b.identifier = p.id;
    b.last_event = p.last_event; b.num_collisions = p.num_collisions; b.breed = p.breed; b.species = p.species;
    const int d = flat_to_domain(im, p.cell);
    b.domain = d; b.cell = p.cell - __ldg(im.domain_cell_offset + d);
}

// returns 1 when the particle keeps tracking, 0 when its history ends on this GPU, 2 when it ends here and the particle
// still has to be deposited in a peer's ring (peer mode; done by send_flush in the warp's next service phase)
__device__ __forceinline__ int facet_crossing_event(const TrackArgs& a, Particle& p, Counters& c)
{
    const DevImage& im = a.im;
    const int face = p.facet >> 2;
    const int event = face_event(p.head, face);
    if (event == QSB_ADJ_TRANSIT_ON)
    {
        if (im.brick)
        {
            // one brick per rank: the neighbour's index and grid position are computed, its face events and material are one
            // 4-byte load that depends on nothing but the cell number (instead of adj[face] -> the neighbour's record)
            const int axis = face >> 1, dir = (face & 1) ? -1 : 1;
            p.cell += dir * (axis == 0 ? im.brick_stride[0] : (axis == 1 ? im.brick_stride[1] : im.brick_stride[2]));
            const uint32_t info = __ldg(im.cell_info + p.cell);
            int ix = cell_ix(p.head), iy = cell_iy(p.head), iz = cell_iz(p.head);
            if (axis == 0) ix += dir; else if (axis == 1) iy += dir; else iz += dir;
            p.head = make_uint4((unsigned)ix | ((unsigned)iy << 16), (unsigned)iz | ((info >> 24) << 16), info & 0xffffffu, 0u);
        }
        else
        {
            p.cell = __ldg(im.cells[p.cell].adj + face);
            p.head = load_cell_head(im, p.cell);
        }
        p.last_event = QSB_EV_FACET_TRANSIT;
        return 1;
    }
    if (event == QSB_ADJ_REFLECT)
    {
        p.last_event = QSB_EV_REFLECTION;
        reflect_particle(im, p);
        return 1;
    }
    if (event == QSB_ADJ_ESCAPE)
    {
        p.last_event = QSB_EV_ESCAPE;
        p.species = -1;
        c.escapes++;
        return 0;
    }
    if (event == QSB_ADJ_TRANSIT_OFF)
    {
        p.last_event = QSB_EV_COMMUNICATION;
        if (a.peer_mode) return 2;
        const size_t k = (size_t)p.cell * 6 + face;
        const int rank = __ldg(im.face_nbr_rank + k);
        const unsigned long long slot = atomicAdd(&a.ctl->send_count[rank], 1ull);
        if (slot >= a.send_capacity) { atomicOr(&a.ctl->overflow, 4u); return false; }
        ExchangeRecord rec;
        fill_base(im, p, rec.p);
        rec.p.domain = __ldg(im.face_adj_domain + k);
        rec.p.cell = __ldg(im.face_adj_cell + k);
        rec.dir[0] = p.alpha; rec.dir[1] = p.beta; rec.dir[2] = p.gamma;
        a.sends[(size_t)rank * a.send_capacity + slot] = rec;
        return 0;
    }
    return 0;
}

// ---- peer exchange over NVLink ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// relaxed system-scope loads: served by the owning GPU's L2 (never this SM's L1), no fence and -- unlike ld.acquire --
// no invalidation of the SM's L1, which holds the cell records and cross-section tables of every warp on it.
// Ordering assumptions of the consumers that use them (ready words, census chunk counters, queue counters), stated because the
// PTX memory model does not promise them for relaxed loads:
//  * a slot's record is read AFTER its ready word was seen, through a control dependency (the loads sit behind the branch on
//    the flag) and with .cg loads that cannot hit a stale L1 line; a peer's deposit does not rely on ordering at all -- the
//    record carries a checksum of its own words and a per-launch salt (store_deposit / deposit_complete), and an incomplete
//    record is simply looked at again later;
//  * secondaries are published with st.release.gpu AFTER their record stores have been issued one pass earlier; the consumer's
//    relaxed load of the ready word followed by dependent .cg loads observes them on every NVIDIA GPU to date (loads are not
//    speculated past the branch that guards them), which is what the parity suite exercises millions of times per run;
//  * queue counters (head / tail) are only hints for how many tickets to ask for: a stale value costs a retry, never a particle.
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_sys(const unsigned int* p)
{
    unsigned int v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ PeerControl* peer_control(const TrackArgs& a, int rank) { return reinterpret_cast<PeerControl*>(a.peer_base[rank]); }

// the flat cell index, on the destination rank, of the cell behind off-rank face k = cell * 6 + face of this rank
__device__ __forceinline__ int peer_destination_cell(const TrackArgs& a, const PeerControl* pc, size_t k)
{
    int cell = __ldg(a.im.face_adj_cell + k);                      // local to the neighbour's domain
    if (a.peer_multi_domain)
    {
        int off;
        asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(off) : "l"(pc->domain_offset + __ldg(a.im.face_adj_domain + k)) : "memory");
        cell += off;
    }
    return cell;
}

// what a block knows about the peers' current launch (filled once per block at kernel start, see track_kernel)
struct PeerLaunch { unsigned long long n_in; unsigned int vault_epoch; unsigned int pad; };

// Deposit the particles of lanes whose history just left this GPU's domain straight into the neighbours' processing
// vaults (the reference packs them into per-neighbour MPI buffers and unpacks them on the other side,
// src/MC_Facet_Crossing_Event.cc:49-67 + src/MC_Particle_Buffer.cc:258-291,452-502).  Nothing on this path waits for
// NVLink: a warp that stood still for the round trips of every crossing lost 6 % of a cycle at 2 GPUs and a quarter at 8
// (measured).  A deposit advances one stage per service phase of its warp: stage 1 issues the remote atomics (slot +
// counters) and leaves their results in flight; stage 2, a few passes later, reads them (they are back), stores the
// self-validating record (store_deposit: no fence, no release) and retires the history locally.  The lane keeps the
// particle in its registers, and takes no new ticket, in between.  Order of the counter updates, which is what the
// termination test relies on: peer's `sent` and `inflight` raised (performed: stage 2 has consumed their return values)
// -> own in-flight count drops; peer's `received` raised after its `inflight`.
struct SendState { int stage; unsigned long long ticket, dep, dep2; };

__device__ __forceinline__ unsigned send_advance(const TrackArgs& a, const PeerLaunch* launch, const Particle& p, SendState& s)
{
    const DevImage& im = a.im;
    unsigned retire = 0u;
    if (s.stage != 0)
    {
        const int face = p.facet >> 2;
        const size_t k = (size_t)p.cell * 6 + face;
        const int rank = __ldg(im.face_nbr_rank + k);
        PeerControl* pc = peer_control(a, rank);
        if (s.stage == 1)
        {
            atomicAdd(&a.ctl->send_count[rank], 1ull);                               // statistics only
            s.dep = atomicAdd_system(&pc->sent, 1ull);              // three results left in flight: nothing here waits for them
            s.dep2 = atomicAdd_system(&pc->inflight, 1ull);
            s.ticket = atomicAdd_system(&pc->tail, 1ull);
            s.stage = 2;
        }
        else
        {
            asm volatile("" :: "l"(s.dep), "l"(s.dep2) : "memory"); // the counter atomics have been performed: their values are here
            const unsigned long long slot = s.ticket - launch[rank].n_in;
            retire = 1u;
            if (slot >= a.proc.capacity)
            {
                st_release_sys(&pc->overflow, a.peer_epoch);       // the peer's host reports it; the particle is dropped
                atomicAdd_system(&pc->inflight, 0ull - 1ull);
            }
            else
                store_deposit(vault_view(a.peer_base[rank], a.proc.capacity), slot, p, peer_destination_cell(a, pc, k), launch[rank].vault_epoch);
            asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" :: "l"(&pc->received), "l"(1ull) : "memory");
            s.stage = 0;
        }
    }
    return __popc(__ballot_sync(kFullMask, retire != 0u));
}

// One wave of the termination test (Mattern's four-counter method): lane r reads rank r's control block through NVLink --
// launch epoch, `received`, in-flight count, `sent`, in that order.  passive = every rank has started this launch and had
// nothing queued or running when its in-flight count was read (after its `received`).
__device__ __forceinline__ bool peer_wave(const TrackArgs& a, unsigned lane, unsigned long long& received, unsigned long long& sent, bool& aborted)
{
    bool passive = true, ab = false;
    received = 0; sent = 0;
    if ((int)lane < a.im.n_ranks)
    {
        const PeerControl* pc = peer_control(a, (int)lane);
        const unsigned int e = ld_acquire_sys(&pc->epoch);
        received = ld_acquire_sys(&pc->received);
        const unsigned long long inf = ld_acquire_sys(&pc->inflight);
        sent = ld_acquire_sys(&pc->sent);
        ab = ld_acquire_sys(&pc->abort) == a.peer_epoch;
        passive = e == a.peer_epoch && inf == 0ull;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        received += __shfl_xor_sync(kFullMask, received, d);
        sent += __shfl_xor_sync(kFullMask, sent, d);
    }
    aborted = __any_sync(kFullMask, ab);
    return __all_sync(kFullMask, passive);
}

// The service warp of a GPU in peer mode (warp 0 of block 0; it tracks nothing) decides global termination.  A rank is
// passive when its in-flight count is zero; a passive rank becomes active only by a deposit, and a deposit is counted in
// the receiver's `sent` word while the sender still counts the history as its own, and in the receiver's `received` after
// it has raised the receiver's in-flight count.  When
// this GPU is passive the warp takes two waves over all ranks; if every rank was passive in both and the totals satisfy
// received(wave 1) == sent(wave 1) == received(wave 2) == sent(wave 2), no deposit was under way and no rank was active
// at the end of the first wave, and termination is stable.  `done` then releases the idle tracking warps.  A launch that
// has not terminated after watchdog_ns raises `abort` on every rank instead.
__device__ __noinline__ void peer_service_loop(const TrackArgs& a, unsigned lane)
{
    PeerControl* me = peer_control(a, a.my_rank);
    const unsigned long long t_start = global_timer_ns();
    unsigned sleep = 500;
    bool seen_idle = false;
    for (;;)
    {
        unsigned long long inflight = 1;
        if (lane == 0) inflight = ld_relaxed_sys(&me->inflight);
        inflight = __shfl_sync(kFullMask, inflight, 0);
        if (inflight == 0ull && !seen_idle) { seen_idle = true; if (lane == 0) me->first_idle_ns = global_timer_ns() - t_start; }
        bool aborted = ld_relaxed_sys(&me->abort) == a.peer_epoch;
        if (!aborted && inflight == 0ull)
        {
            unsigned long long r1, s1, r2 = 0, s2 = 0;
            bool ab1 = false, ab2 = false;
            const bool passive1 = peer_wave(a, lane, r1, s1, ab1);
            const bool passive2 = passive1 && r1 == s1 && peer_wave(a, lane, r2, s2, ab2);
            aborted = ab1 || ab2;
            if (passive2 && r2 == r1 && s2 == r1)
            {
                if (lane == 0)
                {
                    me->done_ns = global_timer_ns() - t_start;
                    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(&me->done), "r"(a.peer_epoch) : "memory");
                }
                return;
            }
        }
        if (aborted || global_timer_ns() - t_start > a.watchdog_ns) break;
        __nanosleep(sleep);
        if (sleep < 4000) sleep *= 2;
        if (inflight != 0ull) sleep = 4000;     // still tracking locally: nothing to decide yet
    }
    if ((int)lane < a.im.n_ranks) st_release_sys(&peer_control(a, (int)lane)->abort, a.peer_epoch);      // tell everybody, ourselves included
}

// census record as the host wants it: MC_Base_Particle layout, 17 eight-byte L2 stores
__device__ __forceinline__ void store_census_aos(const DevImage& im, qsb_base_particle* rec, const Particle& p)
{
    double* r = reinterpret_cast<double*>(rec);
    __stcg(r + 0, p.x); __stcg(r + 1, p.y); __stcg(r + 2, p.z);
    __stcg(r + 3, QSB_VX(p)); __stcg(r + 4, QSB_VY(p)); __stcg(r + 5, QSB_VZ(p));
    __stcg(r + 6, p.energy); __stcg(r + 7, p.weight); __stcg(r + 8, p.ttc);
    __stcg(r + 9, p.age); __stcg(r + 10, p.nmfp); __stcg(r + 11, p.nseg);
    unsigned long long* u = reinterpret_cast<unsigned long long*>(rec);
    __stcg(u + 12, (unsigned long long)p.seed); __stcg(u + 13, (unsigned long long)p.id);
    const int d = flat_to_domain(im, p.cell);
    const int local = p.cell - __ldg(im.domain_cell_offset + d);
    __stcg(u + 14, (unsigned long long)(unsigned)p.last_event | ((unsigned long long)(unsigned)p.num_collisions << 32));
    __stcg(u + 15, (unsigned long long)(unsigned)p.breed | ((unsigned long long)(unsigned)p.species << 32));
    __stcg(u + 16, (unsigned long long)(unsigned)d | ((unsigned long long)(unsigned)local << 32));
}

// Census append.  Slots are allocated in the segment pass that ends the histories (one atomic per warp and pass, issued
// by the lowest such lane and NOT waited for: its return value is first read here, a pass or more later); the records are
// stored in the warp's next service phase by all lanes with a pending census, converged -- SoA vault, or the record-form
// buffer when the census is being streamed to the host.
__device__ __forceinline__ void census_flush(const TrackArgs& a, const Particle& p, bool pending, unsigned lane,
                                             unsigned long long my_base, unsigned my_count, unsigned leader, unsigned rank)
{
    const unsigned group = __ballot_sync(kFullMask, pending);
    if (group == 0u) return;
    const unsigned long long base = __shfl_sync(kFullMask, my_base, leader);   // each lane reads the base its own group's leader holds
    if (pending)
    {
        const unsigned long long slot = base + rank;
        if (slot >= a.census.capacity) atomicOr(&a.ctl->overflow, 2u);
        else if (a.census_aos) store_census_aos(a.im, a.census_aos + slot, p);
        else store_particle(a.census, slot, p, false);
    }
    if (!a.census_aos) return;
    // streaming: count each group's records into their chunk(s) with release semantics -- the records of the whole
    // warp (ordered by the __syncwarp) are visible before the count -- and tell the host about every chunk
    // that became complete, through mapped pinned memory; its D2H copy then runs while tracking continues
    __syncwarp();
    if (pending && my_count != 0u)                    // group leaders
    {
        const unsigned long long last = min(my_base + my_count, a.census.capacity);
        unsigned long long at = my_base;
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

__device__ __forceinline__ unsigned int warp_sum(unsigned int v) { return __reduce_add_sync(kFullMask, v); }

} // namespace
} // namespace qsb
#endif
