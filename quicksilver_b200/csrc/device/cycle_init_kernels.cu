// cycle_init_kernels.cu -- cycleInit on the device: the step directly in front of the tracking
// hot path, src/main.cc:96-121 = MC_SourceNow (src/MC_SourceNow.cc:28-133) + PopulationControl
// (src/PopulationControl.cc:20-122) + RouletteLowWeightParticles (src/PopulationControl.cc:127-171).
//
// The reference runs the three stages one after the other, serially on the host, over an AoS vault that it pops and
// pushes one particle at a time; at benchmark sizes that is ~1 s per cycle next to a tracking kernel of ~20 ms, and it
// forces the whole population through PCIe twice per cycle.  Here the population never leaves HBM: ONE kernel reads last
// cycle's census vault, creates this cycle's source particles in registers, applies population control and the
// low-weight roulette to every particle and appends the survivors and split copies to the processing vault.
//
// Why one pass is the same computation: every decision is taken from the particle's own random-number stream, and the
// three numbers the stages share -- weight of a source particle, split/roulette factor, weight cut-off -- depend only on
// counts known before the kernel starts (census count + source count, summed over ranks by the host model).  Vault
// order is immaterial to the tracker (SURVEY 8a note 9).
//
// Work item i: i < n_carried -> census particle i; otherwise source particle i - n_carried, whose cell is found by
// bisecting the prefix sum of the per-cell source counts.  Output: every thread first counts the records its particle
// will produce, the block reserves that many slots of the processing vault with ONE atomicAdd (10.7 M particles = 42 k
// atomics on the fill counter, not 330 k), and each warp then writes its share "round" by round (split copy k of every
// lane that has one, then the lanes' own particles), each round compacted with a ballot so that a store instruction
// writes consecutive slots of an SoA array.  HBM-bound: reads 164 B and writes 168 B per particle.
//
// Arithmetic: qs_cycle_init.h, the very source the host model runs (pinned byte for byte against the reference by the
// golden fixtures), compiled with --fmad=false and the portable log/sin/cos -> the device's particles equal the host
// model's (strict-math mode) bit for bit.
#include <cuda_runtime.h>
#include <cstdint>

#include "device_types.cuh"
#include "../qs_cycle_init.h"

namespace qsb {
namespace {

constexpr unsigned kFull = 0xffffffffu;

struct InitParticle
{
    double x, y, z, vx, vy, vz, energy, weight, ttc, age, nmfp, nseg;
    unsigned long long seed, id;
    int cell;
    int4 tags;
};

__device__ __forceinline__ void load_census(const VaultView& v, unsigned long long i, InitParticle& p)
{
    p.x = __ldcs(v.x + i); p.y = __ldcs(v.y + i); p.z = __ldcs(v.z + i);
    p.vx = __ldcs(v.vx + i); p.vy = __ldcs(v.vy + i); p.vz = __ldcs(v.vz + i);
    p.energy = __ldcs(v.energy + i); p.weight = __ldcs(v.weight + i); p.ttc = __ldcs(v.ttc + i);
    p.age = __ldcs(v.age + i); p.nmfp = __ldcs(v.nmfp + i); p.nseg = __ldcs(v.nseg + i);
    p.seed = __ldcs(v.seed + i); p.id = __ldcs(v.id + i);
    p.cell = __ldcs(v.cell + i);
    p.tags = __ldcs(v.tags + i);
}

// a record of the processing vault; the direction cosine is left to the tracker (NaN = derive from the velocity, what
// the reference does for every particle that went through MC_Base_Particle, src/MC_Base_Particle.hh:317-325)
__device__ __forceinline__ void store_processing(const VaultView& v, unsigned long long i, const InitParticle& p, double weight,
                                                 unsigned long long seed, unsigned long long id, uint32_t epoch)
{
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    __stcs(v.x + i, p.x); __stcs(v.y + i, p.y); __stcs(v.z + i, p.z);
    __stcs(v.vx + i, p.vx); __stcs(v.vy + i, p.vy); __stcs(v.vz + i, p.vz);
    __stcs(v.energy + i, p.energy); __stcs(v.weight + i, weight); __stcs(v.ttc + i, p.ttc);
    __stcs(v.age + i, p.age); __stcs(v.nmfp + i, p.nmfp); __stcs(v.nseg + i, p.nseg);
    __stcs(v.dirx + i, nan); __stcs(v.diry + i, nan); __stcs(v.dirz + i, nan);
    __stcs(v.seed + i, seed); __stcs(v.id + i, id);
    __stcs(v.cell + i, p.cell);
    __stcs(v.tags + i, p.tags);
    __stcs(v.ready + i, epoch);
}

// one output round of a warp: lanes with `keep` get consecutive slots of the processing vault, starting at `base`
// (the warp's share of the block's reservation), which moves on by the number of records written
__device__ __forceinline__ void emit_round(const CycleInitArgs& a, unsigned lane, bool keep, const InitParticle& p, double weight,
                                           unsigned long long seed, unsigned long long id, unsigned long long& base)
{
    const unsigned mask = __ballot_sync(kFull, keep);
    if (keep)
    {
        const unsigned long long slot = base + __popc(mask & ((1u << lane) - 1u));
        if (slot >= a.dst.capacity) a.out->overflow = 1u;
        else store_processing(a.dst, slot, p, weight, seed, id, a.epoch);
    }
    base += __popc(mask);
}

// The decisions of one particle, taken from its own stream: population control (src/PopulationControl.cc:66-122, not
// entered at all when the factor is exactly 1, :60), then for each split copy -- a copy of the re-weighted particle with
// a stream spawned from the parent's, in order (:106-116) -- and finally for the particle itself the low-weight roulette
// (src/PopulationControl.cc:127-171).  The walk over the copies and the particle is deterministic, so the kernel makes
// it twice: once to count the records it will write, once to write them.
struct Decisions
{
    bool alive;         // survived population control
    int copies;
    uint64_t seed;      // parent's stream after population control
    double weight;      // after population control
};

__device__ __forceinline__ Decisions population_control(const CycleInitArgs& a, bool alive, const InitParticle& p)
{
    Decisions d;
    d.alive = alive; d.copies = 0; d.seed = p.seed; d.weight = p.weight;
    if (alive && a.factor != 1.0)
    {
        const int c = qs_population_control_one(a.factor, &d.seed, &d.weight);
        if (c < 0) d.alive = false; else d.copies = c;
    }
    return d;
}

#ifndef QSB_CI_MIN_BLOCKS
#define QSB_CI_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(256, QSB_CI_MIN_BLOCKS) cycle_init_kernel(const __grid_constant__ CycleInitArgs a)
{
    __shared__ unsigned long long s_warp_base[8];
    __shared__ unsigned s_warp_count[8];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned long long n_total = a.n_carried + a.n_source;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned n_rr = 0, n_split = 0;
    // whole blocks iterate together (the loop bound is rounded up to a multiple of the block), threads past the end stay idle
    const unsigned long long n_round = (n_total + 255ull) & ~255ull;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_round; i += stride)
    {
        const bool present = i < n_total;
        InitParticle p;
        p.weight = 0.0; p.seed = 0; p.id = 0;
        if (present && i < a.n_carried) load_census(a.src, i, p);
        else if (present)
        {
            // source particle j: cell = last cell whose offset is <= j (cells without source particles are skipped)
            const unsigned long long j = i - a.n_carried;
            int lo = 0, hi = a.n_cells;                      // invariant: offsets[lo] <= j < offsets[hi]
            while (hi - lo > 1)
            {
                const int mid = (lo + hi) >> 1;
                if ((unsigned long long)__ldg(a.source_offsets + mid) <= j) lo = mid; else hi = mid;
            }
            const int cell = lo;
            const unsigned long long k = j - (unsigned long long)__ldg(a.source_offsets + cell);
            qs_source_particle s;
            qs_source_one<QsStrictMath>(__ldg(a.source_tally + cell) + k + __ldg(a.cell_id + cell), a.nodes + (size_t)cell * 42,
                                        __ldg(a.cell_volume + cell), a.e_min, a.e_max, a.dt, &s);
            p.x = s.coordinate[0]; p.y = s.coordinate[1]; p.z = s.coordinate[2];
            p.vx = s.velocity[0]; p.vy = s.velocity[1]; p.vz = s.velocity[2];
            p.energy = s.kinetic_energy; p.weight = a.source_weight; p.ttc = s.time_to_census;
            p.age = 0.0; p.nmfp = s.num_mean_free_paths; p.nseg = 0.0;
            p.seed = s.random_number_seed; p.id = s.identifier;
            p.cell = cell;
            p.tags = make_int4(QSB_EV_CENSUS, 0, 0, 0);      // last_event = census (MC_Particle's default), species 0
        }

        const Decisions d = population_control(a, present, p);
        if (present && !d.alive) ++n_rr;
        n_split += (unsigned)d.copies;
        const bool roulette = a.cutoff > 0.0;

        // ---- pass 1: how many records does this particle put into the vault (itself + surviving copies) ----
        unsigned count = 0;
        if (d.alive)
        {
            uint64_t seed = d.seed;
            for (int k = 1; k <= d.copies; ++k)
            {
                uint64_t child_seed = qs_rng_spawn(&seed);
                double child_weight = d.weight;
                if (!roulette || qs_roulette_low_weight_one(a.cutoff, a.weight_cutoff, &child_seed, &child_weight)) ++count; else ++n_rr;
            }
            double weight = d.weight;
            if (!roulette || qs_roulette_low_weight_one(a.cutoff, a.weight_cutoff, &seed, &weight)) ++count; else ++n_rr;
        }
        // ---- one reservation per block: warp totals -> shared memory -> a single atomicAdd on the vault's fill count ----
        const unsigned warp_count = __reduce_add_sync(kFull, count);
        if (lane == 0) s_warp_count[warp] = warp_count;
        __syncthreads();
        if (threadIdx.x == 0)
        {
            unsigned total = 0;
            for (int w = 0; w < 8; ++w) total += s_warp_count[w];
            unsigned long long base = total ? atomicAdd(&a.out->n_out, (unsigned long long)total) : 0ull;
            for (int w = 0; w < 8; ++w) { s_warp_base[w] = base; base += s_warp_count[w]; }
        }
        __syncthreads();
        unsigned long long base = s_warp_base[warp];
        // ---- pass 2: the same walk again, writing.  Round k of a warp = copy k of every lane that has one, then the
        //      lanes' own particles: each round is compacted with a ballot, so a store instruction writes consecutive slots ----
        const int max_copies = __reduce_max_sync(kFull, d.alive ? d.copies : 0);
        uint64_t seed = d.seed;
        for (int k = 1; k <= max_copies; ++k)
        {
            bool keep = d.alive && k <= d.copies;
            uint64_t child_seed = 0, child_id = 0;
            double child_weight = d.weight;
            if (keep)
            {
                child_seed = qs_rng_spawn(&seed);
                child_id = child_seed;
                if (roulette && !qs_roulette_low_weight_one(a.cutoff, a.weight_cutoff, &child_seed, &child_weight)) keep = false;
            }
            emit_round(a, lane, keep, p, child_weight, child_seed, child_id, base);
        }
        bool keep = d.alive;
        double weight = d.weight;
        if (keep && roulette && !qs_roulette_low_weight_one(a.cutoff, a.weight_cutoff, &seed, &weight)) keep = false;
        emit_round(a, lane, keep, p, weight, seed, p.id, base);
        // (s_warp_count / s_warp_base are rewritten only after the next iteration's first barrier / between its two barriers)
    }
    n_rr = __reduce_add_sync(kFull, n_rr);          // per-thread counts are tiny (items per thread x copies)
    n_split = __reduce_add_sync(kFull, n_split);
    if (lane == 0)
    {
        if (n_rr) atomicAdd(&a.out->n_rr, (unsigned long long)n_rr);
        if (n_split) atomicAdd(&a.out->n_split, (unsigned long long)n_split);
    }
}

// the cells' running source counts move on by this cycle's counts (src/MC_SourceNow.cc:92: _sourceTally++ per particle)
__global__ void source_tally_advance_kernel(unsigned long long* __restrict__ tally, const int* __restrict__ offsets, int n_cells)
{
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x)
        tally[c] += (unsigned long long)(offsets[c + 1] - offsets[c]);
}

} // namespace

void launch_cycle_init(const CycleInitArgs& a, int sm_count, cudaStream_t s)
{
    const unsigned long long n_total = a.n_carried + a.n_source;
    if (n_total == 0) return;
    const unsigned long long blocks = (n_total + 255ull) / 256ull;
    const int grid = (int)(blocks < (unsigned long long)sm_count * 8ull ? blocks : (unsigned long long)sm_count * 8ull);
    cycle_init_kernel<<<grid, 256, 0, s>>>(a);
}

void launch_source_tally_advance(unsigned long long* tally, const int* offsets, int n_cells, cudaStream_t s)
{
    if (n_cells <= 0) return;
    const int grid = (n_cells + 255) / 256 < 1184 ? (n_cells + 255) / 256 : 1184;
    source_tally_advance_kernel<<<grid, 256, 0, s>>>(tally, offsets, n_cells);
}

} // namespace qsb
