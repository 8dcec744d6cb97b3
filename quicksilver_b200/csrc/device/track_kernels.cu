// track_kernels.cu -- the cycle-tracking hot path on sm_100a.
//
// One persistent kernel follows particle histories to census / absorption / escape / rank exit
// (reference: CycleTrackingGuts + CycleTrackingFunction, src/CycleTracking.cc:15-119).  Each lane owns one
// in-flight particle held in registers.  Work distribution is a ticket queue over the processing vault:
// an idle lane takes the next ticket (one warp-aggregated atomicAdd, never a CAS loop) and redeems it as
// soon as that slot has been written -- by the host (initial vault, arrivals) or by a fission elsewhere on
// the GPU (secondaries are appended to the same vault).  A single "inflight" counter (histories created
// minus histories finished) tells idle warps when the cycle has drained.  The fissioning parent is
// re-queued in place: the reference sends it through the extra vault and MC_Load_Particle again
// (src/CollisionEvent.cc:137-142); here the same reload transform is applied in registers.
//
// Geometry: the mesh is a uniform brick grid, so the nearest-facet search (24 ray/triangle tests in the
// reference, src/MCT.cc:550-621) is done as a filtered exact predicate: the exit face and the triangle on
// it are found with ordinary box arithmetic and a 1e-6 safety margin, and only the winning facet's
// distance is evaluated with the reference's exact expression on the bit-exact plane rebuilt from the
// 64-byte CellRec.  Anything within the margin of an edge, a diagonal or the exit face itself takes the
// full 24-facet path, which restates the reference statement by statement.  Both paths give the same
// bits; tracking_mode bit 1 makes the kernel run both and count disagreements (tests use it).
//
// Compiled twice (csrc/Makefile): QSB_VALIDATION=1 with --fmad=false and the portable log/sin/cos of
// qs_strict_math.h (bit-comparable with the CPU oracle), QSB_VALIDATION=0 with FMA contraction and the
// CUDA math library.
//
// Reference map: segment outcome  src/MC_Segment_Outcome.cc:31-226
//                nearest facet    src/MCT.cc:87-137,280-395,436-621
//                collision        src/CollisionEvent.cc:25-148, src/NuclearData.cc:54-88,208-227
//                facet crossing   src/MC_Facet_Crossing_Event.cc:25-70, src/MCT.cc:401-429
//                tallies          src/Tallies.hh:36-100,351-354
#include "track_physics.cuh"

namespace qsb {
namespace {


// ---- the persistent history kernel ---------------------------------------------------------------------
// kPeer = 1: the instance launched when the peer exchange is connected.  The single-GPU instance carries none of that
// code: the hot loop is instruction-fetch sensitive (the warp-state samples show it), and the deposit / termination paths
// inlined into the service phase cost 4 % of the single-GPU rate when they were merely present.
template <int kDummy, int kPeer>
#ifndef QSB_MIN_BLOCKS
#define QSB_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(128, QSB_MIN_BLOCKS) track_kernel(const __grid_constant__ TrackArgs a)
{
    const unsigned lane = threadIdx.x & 31u;
    Counters c = {};
    Particle p;
    unsigned long long ticket = kNoTicket;
    unsigned long long pool_next = 0, pool_end = 0;     // warp-uniform: tickets reserved by this warp, not yet handed to a lane
    unsigned long long in_seen = 0;                     // last value of ctl->in_ready this lane has seen
    const uint32_t epoch = a.epoch;
    unsigned backoff = 64;

    bool census_pending = false;                        // history ended at census; the record is still in this lane's registers
    unsigned long long census_base = 0;                 // leader of a census group: first slot of the group (return value of its atomic)
    unsigned census_count = 0, census_leader = 0, census_rank = 0;
    unsigned long long pub_first = 0;                   // secondaries written in the last collision pass, not yet published
    unsigned pub_n = 0;
    unsigned long long spare_base = 0;                  // lane 0: the next ticket batch, reserved one service phase ahead
    bool spare_valid = false;
    int state = kStateIdle;                             // what this lane's particle needs next (kState*)
    unsigned retired = 0;                               // warp-uniform: histories finished since the last service phase
    SendState send = { 0, 0ull, 0ull, 0ull };                 // peer mode: history left for a neighbour's domain, record still in registers
    unsigned long long diag_send_cycles = 0, diag_send_calls = 0;

    // peer mode: before anything can be deposited on a peer, that peer's control words for THIS launch must be in place
    // (its host writes them, then the epoch).  Each block waits for every peer once and keeps what senders need.
    __shared__ PeerLaunch s_launch[kPeer ? kMaxPeers : 1];
    if (kPeer && a.peer_mode)
    {
        if ((int)threadIdx.x < a.im.n_ranks)
        {
            const PeerControl* pc = peer_control(a, (int)threadIdx.x);
            const unsigned long long t0 = global_timer_ns();
            while (ld_acquire_sys(&pc->epoch) != a.peer_epoch && global_timer_ns() - t0 < a.watchdog_ns) __nanosleep(500);
            s_launch[threadIdx.x].n_in = ld_relaxed_sys(&pc->n_in);
            s_launch[threadIdx.x].vault_epoch = ld_relaxed_sys(&pc->vault_epoch);
            if (blockIdx.x == 0 && (int)threadIdx.x != a.my_rank) peer_control(a, a.my_rank)->startup_wait_ns = global_timer_ns() - t0;
        }
        __syncthreads();
        if (blockIdx.x == 0 && threadIdx.x < 32u)
        {
            peer_service_loop(a, lane);                 // this GPU's service warp: global termination, tracks nothing
            return;
        }
    }

    for (;;)
    {
        __syncwarp();
        if (__ballot_sync(kFullMask, pub_n != 0u)) { publish_children(a, pub_first, pub_n, epoch); pub_n = 0u; }

        // SERVICE PHASE.  Ending a history (census store) and starting one (ticket, particle load, reload transform, energy
        // group) are long code paths that single lanes reach at random times; run per lane as they come they would each cost
        // the warp a full pass with one or two lanes active (measured: a third of all issue slots).  Idle lanes therefore wait
        // until kRefill of them have gathered (or nothing is left to track) and then do it together, converged.
        const unsigned idle_mask = __ballot_sync(kFullMask, state == kStateIdle);
        if (__popc(idle_mask) >= kRefill || idle_mask == kFullMask)
        {
            // retire the histories that finished since the last service phase: one reduction per warp.  Secondaries were
            // counted before they became visible, so inflight reaches 0 only when nothing is queued or running.
            if (kPeer && a.peer_mode && __any_sync(kFullMask, send.stage != 0))
            {
                const long long c0 = clock64();
                retired += send_advance(a, s_launch, p, send);
                diag_send_cycles += (unsigned long long)(clock64() - c0); diag_send_calls++;
            }
            if (retired && lane == 0) atomicAdd(a.inflight, 0ull - (unsigned long long)retired);
            retired = 0u;
            census_flush(a, p, census_pending, lane, census_base, census_count, census_leader, census_rank);
            census_pending = false; census_count = 0u;

            // 1. every idle lane without a ticket takes one from the warp's pool.  The pool is refilled kTicketBatch tickets at a
            //    time from a batch that was reserved during the PREVIOUS service phase (one atomicAdd whose return value is
            //    only read now, so its latency is off the critical path).  A reserved ticket is an obligation: it is always
            //    handed to a lane eventually.
            const bool want = state == kStateIdle && ticket == kNoTicket && send.stage == 0;
            const unsigned want_mask = __ballot_sync(kFullMask, want);
            if (want_mask)
            {
                const unsigned long long avail = pool_end - pool_next;
                const unsigned n_want = __popc(want_mask);
                unsigned long long fresh = 0;
                if (n_want > avail)
                {
                    if (!spare_valid && lane == 0) spare_base = atomicAdd(&a.ctl->head, (unsigned long long)kTicketBatch);
                    fresh = __shfl_sync(kFullMask, spare_base, 0);
                    spare_valid = false;
                }
                const unsigned rank = __popc(want_mask & ((1u << lane) - 1u));
                if (want) ticket = rank < avail ? pool_next + rank : fresh + (rank - avail);
                if (n_want > avail) { pool_next = fresh + (n_want - avail); pool_end = fresh + kTicketBatch; if (QSB_OPT_PREFETCH_L2) prefetch_tickets(a, fresh, lane); }
                else pool_next += n_want;
            }
            if (QSB_OPT_SPARE && !spare_valid)
            {
                if (lane == 0) spare_base = atomicAdd(&a.ctl->head, (unsigned long long)kTicketBatch);
                spare_valid = true;
            }
            // 2. redeem.  A streamed host record is there once the DMA front (ctl->in_ready) has passed it; a vault slot is ready
            //    if the host wrote it (below ready_prefix) or its ready word carries this epoch.  The flag is read with a relaxed
            //    gpu-scope load and the particle with L2 (.cg) loads that depend on the branch: the data loads are not issued
            //    before the flag value has come back, and the writer released the flag after the data, so no fence (and no L1
            //    invalidation) is needed on this side.
            bool ready = false;
            if (state == kStateIdle && ticket != kNoTicket)
            {
                if (ticket < a.n_in)
                {
                    if (ticket >= in_seen)
                        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(in_seen) : "l"(&a.ctl->in_ready) : "memory");
                    ready = ticket < in_seen;
                }
                else if (ticket - a.n_in < a.proc.capacity)
                {
                    const unsigned long long slot = ticket - a.n_in;
                    ready = slot < a.ready_prefix;
                    if (!ready)
                    {
                        uint32_t flag;
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(flag) : "l"(a.proc.ready + slot) : "memory");
                        ready = flag == epoch;
                        // a peer's deposit: the flag may have overtaken the record, the record vouches for itself
                        if (__builtin_expect(flag == (epoch | kArrivalBit), 0)) ready = deposit_complete(a.proc, slot, epoch);
                    }
                }
            }
            if (ready)
            {
                if (ticket < a.n_in) { load_particle_aos(a, ticket, p); state = kStateSegment; }
                else state = load_particle(a, ticket - a.n_in, p);
                ticket = kNoTicket;
            }
            if (__ballot_sync(kFullMask, state != kStateIdle) == 0u)
            {
                // idle warp: the cycle is over when no history is queued or running anywhere on this GPU -- or, in peer
                // mode, anywhere on any GPU (the service warp's verdict)
                unsigned long long inflight = 1;
                if (lane == 0)
                {
                    if (!(kPeer && a.peer_mode)) inflight = *((volatile unsigned long long*)a.inflight);
                    else
                    {
                        const PeerControl* me = peer_control(a, a.my_rank);
                        const unsigned done = *((volatile const unsigned int*)&me->done), abrt = *((volatile const unsigned int*)&me->abort);
                        inflight = (done == a.peer_epoch || abrt == a.peer_epoch) ? 0ull : 1ull;
                    }
                }
                inflight = __shfl_sync(kFullMask, inflight, 0);
                if (inflight == 0ull) break;
                __nanosleep(backoff);
                if (backoff < 2048) backoff *= 2;
                continue;
            }
            backoff = 64;
        }

        // TRACKING PHASE, one of two passes per iteration.  A segment ends in a collision about every other time, and the
        // collision code is the longest path of the loop; run straight after each segment it would see half the lanes.  Lanes
        // whose segment ended in a collision therefore wait (kStateCollision), together with freshly loaded secondaries that
        // still need their outgoing trajectory (kStateTail), until kCollide of them have gathered or no lane is left that can
        // advance by a segment; then the warp runs one collision pass for all of them.
        bool finished = false;
        const unsigned seg_mask = __ballot_sync(kFullMask, state == kStateSegment);
        const unsigned col_mask = __ballot_sync(kFullMask, state >= kStateCollision);
        if (__popc(col_mask) >= kCollide || seg_mask == 0u)
        {
            double energy0 = p.energy, angle0 = p.nmfp;             // a raw child carries its sampled outcome in these two fields
            double energy1 = 0.0, angle1 = 0.0, energy2 = 0.0, angle2 = 0.0, energy3 = 0.0, angle3 = 0.0;
            int n_out = 1;
            if (state == kStateCollision)
                n_out = collision_head(a, p, c, energy0, angle0, energy1, angle1, energy2, angle2, energy3, angle3);
            // secondaries of the whole pass: vault slots and the in-flight count with one atomic each per warp
            const unsigned n_child = (state == kStateCollision && n_out > 1) ? (unsigned)(n_out - 1) : 0u;
            unsigned incl = n_child;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
                const unsigned up = __shfl_up_sync(kFullMask, incl, d);
                if ((int)lane >= d) incl += up;
            }
            const unsigned total = __shfl_sync(kFullMask, incl, 31);
            if (total)
            {
                unsigned long long base = 0;
                if (lane == 0)
                {
                    base = atomicAdd(a.tail, (unsigned long long)total);
                    atomicAdd(a.inflight, (unsigned long long)total);
                }
                base = __shfl_sync(kFullMask, base, 0) - a.n_in;     // tickets below n_in are the streamed host records
                if (n_child)
                {
                    const unsigned long long first = base + (incl - n_child);
                    if (first + n_child > a.proc.capacity)
                    {
                        atomicOr(&a.ctl->overflow, 1u);
                        atomicAdd(a.inflight, 0ull - (unsigned long long)n_child);
                    }
                    else
                    {
                        // seeds spawned from the parent's stream in order, before the parent's own trajectory draws
                        // (src/CollisionEvent.cc:125-135)
                        write_raw_child(a, first, p, qs_rng_spawn(&p.seed), energy1, angle1);
                        if (n_child > 1u) write_raw_child(a, first + 1, p, qs_rng_spawn(&p.seed), energy2, angle2);
                        if (n_child > 2u) write_raw_child(a, first + 2, p, qs_rng_spawn(&p.seed), energy3, angle3);
                        pub_first = first; pub_n = n_child;
                    }
                }
            }
            if (state >= kStateCollision)
            {
                if (n_out > 0) { collision_tail(a, p, energy0, angle0, state == kStateTail || n_out > 1); state = kStateSegment; }
                else { state = kStateIdle; finished = true; }
            }
        }
        else
        {
            int outcome = -1;
            if (state == kStateSegment)
            {
                outcome = segment_outcome(a, p, c);
                c.segments++;
                p.nseg += 1.;
                if (outcome == 0) state = kStateCollision;
                else if (outcome == 1)
                {
                    const int go = facet_crossing_event(a, p, c);
                    if (go == 2) { state = kStateIdle; send.stage = 1; }        // retired when the deposit is counted over there
                    else if (go != 1) { state = kStateIdle; finished = true; }
                }
                else { census_pending = true; c.census++; state = kStateIdle; finished = true; }     // stored in the next service phase
            }
            // census slots for the histories that just ended: one atomic per warp, its result is not read before the service phase
            const unsigned cen_mask = __ballot_sync(kFullMask, outcome == 2);
            if (cen_mask)
            {
                const unsigned leader = __ffs(cen_mask) - 1;
                if (outcome == 2)
                {
                    census_leader = leader;
                    census_rank = __popc(cen_mask & ((1u << lane) - 1u));
                    if (lane == leader)
                    {
                        census_count = __popc(cen_mask);
                        census_base = atomicAdd(&a.ctl->census_count, (unsigned long long)census_count);
                    }
                }
            }
        }
        // finished histories are retired (subtracted from the global in-flight count) in the next service phase
        retired += __popc(__ballot_sync(kFullMask, finished));
    }

    // flush the per-thread balance counters: warp sum, one atomic per counter per warp
    __syncwarp();
    if (kPeer && a.peer_mode && lane == 0 && diag_send_calls)
    {
        atomicAdd(&peer_control(a, a.my_rank)->send_cycles, diag_send_cycles);
        atomicAdd(&peer_control(a, a.my_rank)->send_calls, diag_send_calls);
    }
    const unsigned int s_seg = warp_sum(c.segments), s_col = warp_sum(c.collisions);
    const unsigned int s_abs = warp_sum(c.absorbs), s_fis = warp_sum(c.fissions), s_pro = warp_sum(c.produced);
    const unsigned int s_esc = warp_sum(c.escapes), s_cen = warp_sum(c.census);
#if QSB_VALIDATION
    const unsigned int s_sca = warp_sum(c.scatters), s_look = warp_sum(c.lookups), s_slow = warp_sum(c.slow), s_mis = warp_sum(c.mismatch);
#else
    // every collision of the fast build is one of the three reaction types (an undefined type would have been rejected by the host model)
    const unsigned int s_sca = s_col - s_abs - s_fis, s_look = 0u, s_slow = 0u, s_mis = 0u;
#endif
    if (lane == 0)
    {
        unsigned long long* b = a.ctl->balance;
        if (s_seg) atomicAdd(b + QSB_BAL_NUM_SEGMENTS, (unsigned long long)s_seg);
        if (s_col) atomicAdd(b + QSB_BAL_COLLISION, (unsigned long long)s_col);
        if (s_sca) atomicAdd(b + QSB_BAL_SCATTER, (unsigned long long)s_sca);
        if (s_abs) atomicAdd(b + QSB_BAL_ABSORB, (unsigned long long)s_abs);
        if (s_fis) atomicAdd(b + QSB_BAL_FISSION, (unsigned long long)s_fis);
        if (s_pro) atomicAdd(b + QSB_BAL_PRODUCE, (unsigned long long)s_pro);
        if (s_esc) atomicAdd(b + QSB_BAL_ESCAPE, (unsigned long long)s_esc);
        if (s_cen) atomicAdd(b + QSB_BAL_CENSUS, (unsigned long long)s_cen);
        if (s_look) atomicAdd(&a.ctl->n_lookups, (unsigned long long)s_look);
        if (s_slow) atomicAdd(&a.ctl->slow_geometry, (unsigned long long)s_slow);
        if (s_mis) atomicAdd(&a.ctl->geometry_mismatch, (unsigned long long)s_mis);
    }
}

} // namespace

#if QSB_VALIDATION
#define QSB_LAUNCH_NAME launch_track_validation
#define QSB_ATTR_NAME track_kernel_attributes_validation
#else
#define QSB_LAUNCH_NAME launch_track_fast
#define QSB_ATTR_NAME track_kernel_attributes_fast
#endif

void QSB_LAUNCH_NAME(const TrackArgs& a, int grid, int block, cudaStream_t s)
{
    if (a.peer_mode) track_kernel<QSB_VALIDATION, 1><<<grid, block, 0, s>>>(a);
    else             track_kernel<QSB_VALIDATION, 0><<<grid, block, 0, s>>>(a);
}

void QSB_ATTR_NAME(int* regs, int* max_blocks_per_sm, int block)
{
    cudaFuncAttributes attr;
    // both instances must be fully resident (the kernel is persistent): report the tighter of the two
    int r = 0, nb = 1 << 30;
    if (cudaFuncGetAttributes(&attr, track_kernel<QSB_VALIDATION, 0>) == cudaSuccess) r = attr.numRegs;
    if (cudaFuncGetAttributes(&attr, track_kernel<QSB_VALIDATION, 1>) == cudaSuccess && attr.numRegs > r) r = attr.numRegs;
    if (regs) *regs = r;
    int n0 = 0, n1 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n0, track_kernel<QSB_VALIDATION, 0>, block, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n1, track_kernel<QSB_VALIDATION, 1>, block, 0);
    nb = n0 < n1 ? n0 : n1;
    if (max_blocks_per_sm) *max_blocks_per_sm = nb;
}

} // namespace qsb
