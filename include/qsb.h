/* qsb.h -- C ABI of the B200-native cycle-tracking path for Quicksilver-class Monte Carlo transport.
 *
 * The reference (LLNL/Quicksilver) has no plugin/FFI interface; its de-facto boundary is
 *   void CycleTrackingGuts(MonteCarlo*, int particle_index, ParticleVault* processing, ParticleVault* processed)
 *   (src/CycleTracking.hh:8-14), selected by the ExecutionPolicy switch inside
 *   cycleTracking(MonteCarlo*) (src/main.cc:138-307, src/cudaUtils.hh:10-25).
 * This header is what a maintainer would bind in its place: plain pointers and sizes, no C++ or torch
 * types, every call returns 0 or a negative qsb_status and never throws or aborts (the reference's own
 * convention is "print and continue", src/qs_assert.hh:9-26; here the message is kept in
 * qsb_last_error()).  INTEGRATION.md shows the binding.
 *
 * Three groups of entry points:
 *   qsb_mc_*    host model: the reference's Parameters / initMC / cycleInit / cycleFinalize surface
 *               (deck + CLI parsing, mesh + nuclear-data construction, source, population control,
 *               balance bookkeeping).  Pure host code, usable without a GPU.
 *   qsb_*       device context: one per GPU; uploads the flattened problem image, owns the SoA particle
 *               vaults, runs the tracking kernels.  This is the hot path.  There is NO CPU fallback:
 *               every call fails with QSB_ERR_CUDA when no sm_100 device is usable.
 *   qsb_mc_cycle_tracking   the drop-in for the reference's cycleTracking(): host vault -> device,
 *               track to exhaustion, census + tallies back to the host model.
 */
#ifndef QSB_H
#define QSB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QSB_ABI_VERSION 1

typedef enum qsb_status {
    QSB_OK            =  0,
    QSB_ERR_ARG       = -1,   /* bad argument / null pointer                      */
    QSB_ERR_INPUT     = -2,   /* deck or command line could not be understood      */
    QSB_ERR_CUDA      = -3,   /* CUDA runtime error or no usable device            */
    QSB_ERR_CAPACITY  = -4,   /* a fixed-capacity vault or slab overflowed         */
    QSB_ERR_STATE     = -5,   /* call made in the wrong phase of a cycle           */
    QSB_ERR_INTERNAL  = -6
} qsb_status;

/* ---- particle record: byte-for-byte the reference's MC_Base_Particle (src/MC_Base_Particle.hh:75-92,
 *      136 bytes: 12 f64, 2 u64, 6 i32) so host vaults can be handed over without conversion. ---- */
typedef struct qsb_base_particle {
    double   coordinate[3];
    double   velocity[3];
    double   kinetic_energy;
    double   weight;
    double   time_to_census;
    double   age;
    double   num_mean_free_paths;
    double   num_segments;
    uint64_t random_number_seed;
    uint64_t identifier;
    int32_t  last_event;        /* MC_Tally_Event (src/Tallies.hh:24-35)            */
    int32_t  num_collisions;
    int32_t  breed;
    int32_t  species;           /* -1 = invalid slot                                 */
    int32_t  domain;            /* rank-local domain index                           */
    int32_t  cell;              /* domain-local cell index                           */
} qsb_base_particle;

/* balance counters in the order the reference reduces them (src/Tallies.cc:31-43) */
enum { QSB_BAL_ABSORB = 0, QSB_BAL_CENSUS, QSB_BAL_ESCAPE, QSB_BAL_COLLISION, QSB_BAL_END, QSB_BAL_FISSION,
       QSB_BAL_PRODUCE, QSB_BAL_SCATTER, QSB_BAL_START, QSB_BAL_SOURCE, QSB_BAL_RR, QSB_BAL_SPLIT,
       QSB_BAL_NUM_SEGMENTS, QSB_BAL_COUNT };

/* facet adjacency events (src/MC_Facet_Adjacency.hh:9-20) */
enum { QSB_ADJ_UNDEFINED = 0, QSB_ADJ_ESCAPE = 1, QSB_ADJ_REFLECT = 2, QSB_ADJ_TRANSIT_ON = 3, QSB_ADJ_TRANSIT_OFF = 4 };

/* tally events (src/Tallies.hh:24-35) */
enum { QSB_EV_COLLISION = 0, QSB_EV_FACET_TRANSIT = 1, QSB_EV_CENSUS = 2, QSB_EV_TRACKING_ERROR = 3,
       QSB_EV_ESCAPE = 4, QSB_EV_REFLECTION = 5, QSB_EV_COMMUNICATION = 6 };

/* reaction types (src/NuclearData.hh:33-39) */
enum { QSB_REACT_UNDEFINED = 0, QSB_REACT_SCATTER = 1, QSB_REACT_ABSORPTION = 2, QSB_REACT_FISSION = 3 };

/* ---- flattened, read-only problem image: what the tracking path reads.  All the rank's domains are
 *      concatenated into one cell index space ("flat cell" = domain_cell_offset[domain] + cell).
 *      Every value is bit-identical to what the reference computes for the same deck (pinned by
 *      oracle/ref_dump.cc).  Pointers are borrowed from the qsb_mc that built them. ---- */
typedef struct qsb_image {
    int32_t  abi_version;
    int32_t  n_domains;              /* domains held by this rank (src/initMC.cc:256-259)            */
    int32_t  n_cells;                /* over all local domains                                        */
    int32_t  n_groups;
    int32_t  n_materials;
    int32_t  n_isotopes;             /* over all materials                                            */
    int32_t  max_reactions_per_material; /* max over materials of nIsotopes*nReactions                */
    int32_t  my_rank, n_ranks;
    int32_t  global_nx, global_ny, global_nz;
    double   global_lx, global_ly, global_lz;
    const int32_t*  domain_cell_offset;  /* [n_domains+1]                                             */
    const int32_t*  domain_gid;          /* [n_domains] global domain id                              */
    /* geometry, per flat cell */
    const double*   planes;          /* [n_cells][24][4] A,B,C,D  (src/MC_Facet_Geometry.hh:18-39)     */
    const double*   nodes;           /* [n_cells][14][3] the cell's own 14 points (src/GlobalFccGrid.cc:48-70) */
    const int32_t*  cell_gid;        /* [n_cells] global cell id ix + nx*(iy + ny*iz)                 */
    const int32_t*  cell_material;   /* [n_cells]                                                     */
    const double*   cell_volume;     /* [n_cells]                                                     */
    const uint64_t* cell_id;         /* [n_cells] seed base, gid << 32 (src/MC_Domain.cc:390)         */
    /* adjacency, per flat cell and FACE (the 4 facets of a face share it: src/MC_Domain.cc:292-318) */
    const uint8_t*  face_event;      /* [n_cells][6] QSB_ADJ_*                                        */
    const int32_t*  face_adj_cell;   /* [n_cells][6] on-processor: flat cell; off-processor: cell index
                                        local to the neighbour's domain; boundary: own flat cell      */
    const int32_t*  face_adj_domain; /* [n_cells][6] neighbour's rank-local domain index              */
    const int32_t*  face_nbr_rank;   /* [n_cells][6] owning rank for off-processor faces, else -1     */
    /* nuclear data */
    const double*   energies;        /* [n_groups+1] group edges (src/NuclearData.cc:105-119)         */
    const int32_t*  mat_n_isotopes;  /* [n_materials]                                                 */
    const int32_t*  mat_n_reactions; /* [n_materials] reactions per isotope                           */
    const double*   mat_mass;        /* [n_materials]                                                 */
    const double*   mat_nu_bar;      /* [n_materials] nuBar of the material's reactions               */
    const uint8_t*  mat_react_type;  /* [n_materials][max_reactions_per_material] QSB_REACT_* by (iso,react) */
    const double*   xs_total;        /* [n_materials][n_groups] = weightedMacroscopicCrossSection
                                        (src/MacroscopicCrossSection.cc:59-80); identical for every
                                        cell of a material because cellNumberDensity == 1
                                        (src/MC_Domain.cc:387)                                        */
    const double*   xs_react;        /* [n_materials][n_groups][max_reactions_per_material]
                                        atomFraction*density*sigma in (iso,react) scan order
                                        (src/CollisionEvent.cc:67-83)                                 */
    const uint8_t*  mat_periodic;    /* [n_materials] 1 if every isotope of the material carries the
                                        same reaction table (always true for reference decks)        */
} qsb_image;

/* ======================================================================================================
 * Host model (qsb_mc_*)
 * ==================================================================================================== */
typedef struct qsb_mc qsb_mc;

/* reduction over ranks used by cycleInit/cycleFinalize and the benchmark report, the stand-in for mpiAllreduce
 * (src/utilsMpi.hh:24-50).  dtype: 0 = f64 sum, 1 = u64 sum, 2 = f64 max.  In place.  NULL (default) = single rank. */
typedef void (*qsb_allreduce_fn)(void* user, void* buf, int32_t count, int32_t dtype);

/* Parse argv exactly like the reference's getParameters (CLI first, deck overrides: src/Parameters.cc:80-95)
 * and build the rank's MonteCarlo model (src/initMC.cc:53-74).  rank/n_ranks replace MPI_Comm_rank/size. */
int  qsb_mc_create(int argc, const char* const* argv, int rank, int n_ranks, qsb_mc** out);
int  qsb_mc_destroy(qsb_mc* mc);
int  qsb_mc_set_allreduce(qsb_mc* mc, qsb_allreduce_fn fn, void* user);
/* echo of the parameters in deck syntax ("output is a valid input", src/Parameters.cc:97-215). */
int  qsb_mc_print_parameters(qsb_mc* mc, char* buf, uint64_t cap, uint64_t* needed);
int  qsb_mc_get_image(qsb_mc* mc, qsb_image* out);
int  qsb_mc_get_int(qsb_mc* mc, const char* key, int64_t* out);       /* nSteps, nParticles, nx, xDom, ... */
int  qsb_mc_get_double(qsb_mc* mc, const char* key, double* out);     /* dt, lx, eMin, source_particle_weight ... */

/* cycleInit (src/main.cc:96-121): swap census -> processing, source, population control, roulette. */
int  qsb_mc_cycle_init(qsb_mc* mc);
/* 0 (default): MC_SourceNow evaluates log/sin/cos with libm -- the reference binary's bits; 1: with the portable
 * functions of csrc/qs_strict_math.h, which the device evaluates to the same bits (validation of the device cycleInit). */
int  qsb_mc_set_strict_math(qsb_mc* mc, int on);
/* processing vault (tracking input) as one contiguous AoS; valid until the next qsb_mc_* call. */
int  qsb_mc_processing(qsb_mc* mc, const qsb_base_particle** aos, uint64_t* n);
/* processed vault (this cycle's census, the next cycle's carried-over particles); n may be NULL. */
int  qsb_mc_processed(qsb_mc* mc, const qsb_base_particle** aos, uint64_t* n);
/* hand the tracker's result back: census particles become the processed vault, counters are added to
 * this cycle's balance task (only the tracking counters are read: absorb, census, escape, collision,
 * fission, produce, scatter, num_segments). */
int  qsb_mc_set_tracking_result(qsb_mc* mc, const qsb_base_particle* census, uint64_t n_census,
                                const uint64_t balance[QSB_BAL_COUNT], double scalar_flux_sum);
/* cycleFinalize (src/main.cc:310-324, src/Tallies.cc:25-98): reduce, accumulate, roll _start.
 * row[0..12] = this cycle's global balance in QSB_BAL_* order, *flux = global scalar-flux sum. */
int  qsb_mc_cycle_finalize(qsb_mc* mc, uint64_t row[QSB_BAL_COUNT], double* flux);
int  qsb_mc_cumulative_balance(qsb_mc* mc, uint64_t out[QSB_BAL_COUNT]);
/* EnergySpectrum (src/EnergySpectrum.cc:12-62; kept only when the deck or `-e` names a spectrum file): per group edge,
 * the number of census particles the cycles so far left in that group (every cycle's census counts,
 * Tallies::CycleFinalize src/Tallies.cc:97).  qsb_mc_energy_spectrum: n_groups+1 counts summed over ranks (every rank
 * calls it); qsb_mc_write_energy_spectrum = PrintSpectrum: rank 0 writes <name>.dat, "index\tenergy\tcount" per line.
 * (The reference allocates n_groups counters but reduces and prints n_groups + 1, src/MonteCarlo.cc:39-48 vs
 * src/EnergySpectrum.cc:41-56: its last line is a read past the end.  Here the last counter is what the index means,
 * particles above eMax -- always 0 for the reference decks; all other lines equal the reference's.) */
int  qsb_mc_energy_spectrum(qsb_mc* mc, uint64_t* counts, uint64_t cap, uint64_t* n);
int  qsb_mc_write_energy_spectrum(qsb_mc* mc);
/* checkCrossSections (src/initMC.cc:392-484): per group the absorption / fission / scatter cross section of every material;
 * written to <crossSectionsOut>.dat by qsb_mc_create when the deck or `-S` names a file; this returns the same text. */
int  qsb_mc_cross_sections_text(qsb_mc* mc, char* buf, uint64_t cap, uint64_t* needed);
/* one line of the reference's cycle table (src/Tallies.cc:123-144, src/Tallies.hh:60-76). */
/* coralBenchmarkCorrectness (src/CoralBenchmark.cc:17-226): the reference's end-of-run self checks for the CORAL decks --
 * reaction ratios, collisions vs facet crossings, lost particles (all from the cumulative balance) and fluence homogeneity
 * over this rank's cells (max over ranks through the allreduce hook) -- printed with the reference's own wording.  Empty
 * when the deck does not set coralBenchmark.  Every rank calls it (the fluence test reduces); rank 0's text is the report.
 * *passed (optional): number of tests that passed, out of 4. */
int  qsb_mc_coral_benchmark_report(qsb_mc* mc, const double* fluence, uint64_t n_cells, char* buf, uint64_t cap, uint64_t* needed,
                                   int32_t* passed);
/* The reference's timer table (MC_Fast_Timer, src/MC_Fast_Timer.{hh,cc}): seven named wall-clock sections in microseconds.
 * The library times the qsb_mc_* calls that cover a whole section -- cycleInit (qsb_mc_cycle_init[_resident]), cycleTracking
 * (qsb_mc_cycle_tracking[_resident], or qsb_mc_tracking_begin ... qsb_mc_tracking_end, or the end of
 * qsb_mc_cycle_init_resident ... qsb_mc_tracking_end_resident), cycleTracking_Kernel (CUDA-event time of the launches),
 * cycleFinalize, main (since qsb_mc_create) -- and the caller adds what it runs itself (a multi-rank driver's exchange:
 * cycleTracking_MPI, cycleTracking_Test_Done) with qsb_mc_timer_add.
 * qsb_mc_format_timer_report: last_cycle = 0 -> Cumulative_Report (src/MC_Fast_Timer.cc:58-105): heading, one line per timer
 * (calls, min / avg / max / stddev over ranks, efficiency) and the Figure Of Merit line = segments / max over ranks of the
 * cycleTracking time; last_cycle = 1 -> Last_Cycle_Report (:107-152, what `cycleTimers: 1` prints after every cycle).
 * Every rank calls it (it reduces); the text is rank 0's. */
enum { QSB_TIMER_MAIN = 0, QSB_TIMER_CYCLE_INIT, QSB_TIMER_CYCLE_TRACKING, QSB_TIMER_CYCLE_TRACKING_KERNEL, QSB_TIMER_CYCLE_TRACKING_MPI,
       QSB_TIMER_CYCLE_TRACKING_TEST_DONE, QSB_TIMER_CYCLE_FINALIZE, QSB_TIMER_COUNT };
int  qsb_mc_timer_add(qsb_mc* mc, int timer, double microseconds, uint64_t calls);
int  qsb_mc_get_timer(qsb_mc* mc, int timer, double* cumulative_microseconds, uint64_t* calls);
int  qsb_mc_format_timer_report(qsb_mc* mc, int last_cycle, char* buf, uint64_t cap, uint64_t* needed);
/* the reference's closing line: Figure Of Merit = segments / cycle-tracking seconds (src/MC_Fast_Timer.cc:97-104) */
int  qsb_mc_format_figure_of_merit(qsb_mc* mc, double tracking_seconds, char* buf, uint64_t cap);
int  qsb_mc_format_cycle_row(qsb_mc* mc, int cycle, const uint64_t row[QSB_BAL_COUNT], double flux,
                             double t_init, double t_track, double t_final, char* buf, uint64_t cap);
const char* qsb_mc_last_error(qsb_mc* mc);

/* ======================================================================================================
 * Device context (qsb_*) -- the hot path
 * ==================================================================================================== */
typedef struct qsb_ctx qsb_ctx;

typedef struct qsb_options {
    int32_t  validation;        /* 1: --fmad=false kernels + strict log/sin/cos (bit-exact vs oracle);
                                   0: fast build (FMA contraction of the same source)                  */
    int32_t  tracking_mode;     /* bit 0: 0 = event-based kernel (the default: particles in shared memory, every warp runs
                                   batches of ONE event type over its own slots), 1 = history-based persistent kernel (one
                                   lane = one history in registers); same results bit for bit; QSB_TRACKING=history|event
                                   in the environment overrides.  bit 1 (value 2): run the
                                   filtered and the full nearest-facet search side by side and count mismatches;
                                   bit 2 (value 4): likewise for the direct reaction selection vs the subtraction chain */
    uint64_t particle_capacity; /* SoA slots per vault; 0 = 1 << 20 (callers size it: nParticles x (3 + 2 nuBar)) */
    uint64_t send_capacity;     /* slots per peer send/recv slab; 0 = derive                             */
    int32_t  threads_per_block; /* 0 = default                                                           */
    int32_t  blocks_per_sm;     /* 0 = default                                                           */
} qsb_options;

typedef struct qsb_track_stats {
    uint64_t n_processed;       /* particle records consumed from the processing vault (incl. secondaries) */
    uint64_t n_census;          /* records now in the census vault                                      */
    uint64_t n_sent;            /* records written to peer send slabs                                   */
    uint32_t n_launches;        /* kernels launched by this call                                        */
    float    device_ms;         /* CUDA-event time of the call's stream work                            */
} qsb_track_stats;

int  qsb_create(int device, const qsb_image* image, double time_step, const qsb_options* opt, qsb_ctx** out);
int  qsb_destroy(qsb_ctx* ctx);
/* clearCrossSectionCache + per-cycle flux/balance reset (src/main.cc:101-103, src/Tallies.cc:75-93);
 * swaps census -> processing on the device when keep_census != 0. */
int  qsb_cycle_begin(qsb_ctx* ctx, int keep_census);
/* ---- device-resident cycles: cycleInit on the device (src/main.cc:96-121), the population never leaves HBM -------------
 * qsb_cycle_init_resident stands for qsb_cycle_begin + qsb_put_particles of a cycle: it clears the cycle's tallies like
 * qsb_cycle_begin and then fills the processing vault ON THE DEVICE from (i) the census vault of the previous cycle and
 * (ii) this cycle's source particles (MC_SourceNow, src/MC_SourceNow.cc:28-133), each passed through population control
 * (src/PopulationControl.cc:20-122) and the low-weight roulette (src/PopulationControl.cc:127-171) -- one kernel, same
 * per-particle random-number streams and arithmetic as the host model (qsb_mc_cycle_init in strict-math mode gives the
 * same particles bit for bit).  The caller supplies the numbers that need other ranks: the weight of a source particle, the
 * per-cell source counts that follow from it, and the split / roulette factor target / global count.  qsb_track follows
 * as usual; the census stays in the census vault for the next qsb_cycle_init_resident (or qsb_get_census).
 * Not available while the previous cycle's census was streamed to the host (qsb_stream_begin): that census is not in
 * the vault.  qsb_put_census seeds the census vault from host records (first resident cycle after host cycles, restart). */
typedef struct qsb_cycle_init_args {
    uint64_t        plan_id;            /* the two arrays below are read (and kept on the device) only when this differs
                                           from the previous call's; the device advances its running counts itself     */
    const int32_t*  source_offsets;     /* [n_cells+1] prefix sum over flat cells of (int)(cellWeight / source_weight)
                                           (src/MC_SourceNow.cc:72-76)                                                  */
    const uint64_t* source_tally;       /* [n_cells] the cells' running source counts before this cycle
                                           (MC_Cell_State::_sourceTally, src/MC_SourceNow.cc:92)                       */
    double          source_weight;      /* weight of one source particle (src/MC_SourceNow.cc:59-61)                   */
    double          e_min, e_max;       /* source energy range (src/MC_SourceNow.cc:108-110)                           */
    double          split_factor;       /* population-control factor (src/PopulationControl.cc:32-57); 1.0 = none       */
    double          low_weight_cutoff;  /* deck lowWeightCutoff, relative to source_weight; <= 0 = off                 */
} qsb_cycle_init_args;

typedef struct qsb_cycle_init_result {
    uint64_t n_start;                   /* census particles carried over (Balance::_start)                             */
    uint64_t n_source;                  /* source particles created (Balance::_source)                                 */
    uint64_t n_rr;                      /* killed by population control + low-weight roulette (Balance::_rr)           */
    uint64_t n_split;                   /* split copies made (Balance::_split)                                         */
    uint64_t n_processing;              /* records now in the processing vault                                         */
    float    device_ms;                 /* CUDA-event time of the kernel                                               */
    uint32_t n_launches;
} qsb_cycle_init_result;

int  qsb_cycle_init_resident(qsb_ctx* ctx, const qsb_cycle_init_args* args, qsb_cycle_init_result* result);
/* host records -> census vault (replaces its contents); the next qsb_cycle_init_resident carries them over. */
int  qsb_put_census(qsb_ctx* ctx, const qsb_base_particle* aos, uint64_t n);

/* host AoS vault -> device SoA processing vault (append). */
int  qsb_put_particles(qsb_ctx* ctx, const qsb_base_particle* aos, uint64_t n);
/* run all local histories to exhaustion, secondaries included (src/main.cc:163-283 for one rank). */
int  qsb_track(qsb_ctx* ctx, qsb_track_stats* stats);
/* Host-buffer streaming: the same cycle with the copies overlapped with tracking.  qsb_stream_begin (directly after
 * qsb_cycle_begin) names the cycle's input vault `in` and the buffer the census is to be delivered to; the next qsb_track
 * launches the tracking kernel first and then feeds it: the host vault is DMA-copied in 8.9 MB chunks as it is (136-byte
 * records, read directly by the kernel) while finished census chunks travel back the same way.  Buffers should be
 * page-locked (the host model's vaults are); pageable ones are staged through an internal bounce buffer.  Further
 * qsb_track calls of the same cycle (arrivals from other ranks) keep appending to the census; qsb_stream_end copies what is
 * left and returns the census count.  If it exceeds census_cap the first census_cap records were delivered and the rest
 * can be fetched with qsb_get_census_range.  qsb_track_host = begin + track + end (one rank). */
int  qsb_stream_begin(qsb_ctx* ctx, const qsb_base_particle* in, uint64_t n_in, qsb_base_particle* census_out, uint64_t census_cap);
int  qsb_stream_end(qsb_ctx* ctx, uint64_t* n_census);
int  qsb_get_census_range(qsb_ctx* ctx, uint64_t first, qsb_base_particle* out, uint64_t count);
int  qsb_track_host(qsb_ctx* ctx, const qsb_base_particle* in, uint64_t n_in, qsb_base_particle* census_out, uint64_t census_cap,
                    uint64_t* n_census, qsb_track_stats* stats);
int  qsb_census_count(qsb_ctx* ctx, uint64_t* n);
int  qsb_get_census(qsb_ctx* ctx, qsb_base_particle* aos, uint64_t cap, uint64_t* n);
int  qsb_get_balance(qsb_ctx* ctx, uint64_t out[QSB_BAL_COUNT]);
int  qsb_get_scalar_flux(qsb_ctx* ctx, double* out /* [n_cells][n_groups] */);
int  qsb_scalar_flux_sum(qsb_ctx* ctx, double* sum);
/* Fluence (src/Tallies.cc:100-121, part of Tallies::CycleFinalize for the CORAL benchmark decks): add this cycle's scalar
 * flux, summed over groups, to the running per-cell fluence -- on the device, before the next qsb_cycle_begin clears the
 * flux; qsb_get_fluence copies the n_cells running totals to the host. */
int  qsb_fluence_accumulate(qsb_ctx* ctx);
/* EnergySpectrum::UpdateSpectrum (src/EnergySpectrum.cc:12-35) for a census that stays on the device: counts[g] = number
 * of records of the current census whose kinetic energy lies in group g (NuclearData::getEnergyGroup,
 * src/NuclearData.cc:208-227); n_counts must be n_groups + 1.  One histogram kernel over the census energies. */
int  qsb_census_energy_spectrum(qsb_ctx* ctx, uint64_t* counts, uint64_t n_counts);
int  qsb_get_fluence(qsb_ctx* ctx, double* out /* [n_cells] */);
/* boundary-particle exchange (src/MC_Facet_Crossing_Event.cc:49-67, src/MC_Particle_Buffer.cc): the
 * tracker packs leavers into per-peer slabs of qsb_base_particle + direction cosine (160 B records);
 * the caller moves slabs between ranks (NCCL send/recv) and feeds arrivals back with qsb_put_arrivals. */
int  qsb_send_counts(qsb_ctx* ctx, uint64_t* counts /* [n_ranks] */);
int  qsb_send_slab(qsb_ctx* ctx, int peer, void** device_ptr, uint64_t* n_records);
int  qsb_clear_sends(qsb_ctx* ctx);
int  qsb_put_arrivals(qsb_ctx* ctx, const void* device_records, uint64_t n_records);
uint64_t qsb_exchange_record_bytes(void);
/* Peer exchange over NVLink / NVSwitch (replaces the per-round pack / MPI_Isend / MPI_Irecv / unpack protocol of
 * src/MC_Particle_Buffer.cc:176-291,452-618 when all ranks drive GPUs of one node, at most QSB_MAX_PEERS of them, one domain
 * per rank).  Each rank exports one device allocation -- a control block followed by its processing vault -- as a 64-byte
 * CUDA IPC handle; after qsb_peer_connect with every rank's handle the tracking kernel stores a boundary-crossing particle
 * straight into a slot of the neighbour's processing vault, where the neighbour's kernel finds it in its ticket queue while
 * it is still tracking, and global termination (the reference's allreduce of sends/receives,
 * src/MC_Particle_Buffer.cc:601-618) is decided on the devices.  A cycle is then ONE qsb_track call per rank, no rounds.
 * Contract: all ranks use the same particle_capacity (qsb_peer_export returns it for the caller to compare) and call
 * qsb_track the same number of times, with a cross-rank synchronisation point between two peer-mode launches: a rank must
 * not start launch k+1 before every rank has returned from launch k (the termination waves compare every rank's launch
 * epoch; a rank that ran ahead would keep a slower one waiting for its watchdog).  The reference's cycle has such a point
 * for free -- cycleFinalize's allreduce of the balance tallies, src/Tallies.cc:75-93 via src/main.cc:310-324 -- and so do
 * quicksilver_b200.driver and the qs_b200 executable; a caller that launches back to back must add its own barrier.
 * watchdog_seconds (0 = 60): a launch that has not terminated by then is abandoned on every rank and qsb_track fails. */
#define QSB_MAX_PEERS 8
#define QSB_PEER_HANDLE_BYTES 64
int  qsb_peer_export(qsb_ctx* ctx, void* handle /* [QSB_PEER_HANDLE_BYTES] */, uint64_t* vault_capacity);
int  qsb_peer_connect(qsb_ctx* ctx, const void* handles /* [n_ranks][QSB_PEER_HANDLE_BYTES], rank order */, int n_ranks,
                      double watchdog_seconds);
int  qsb_peer_disconnect(qsb_ctx* ctx);
/* timings of the last peer-mode qsb_track on this rank, measured on the device: [0] kernel start -> this GPU first had nothing
 * queued or running (ns), [1] kernel start -> global termination seen (ns), [2] SM cycles its warps spent depositing particles
 * on peers (summed over warps), [3] deposit passes, [4] wait for the peers' launch at kernel start (ns), [5] tickets used,
 * [6] particles deposited on peers, [7] kernel start -> a warp first found this GPU's own vault queue empty (ns; event kernel:
 * what is tracked after that are arrivals and the chains of hops they start).  The tail of a cycle is [1] - [0]. */
int  qsb_peer_diagnostics(qsb_ctx* ctx, uint64_t out[8]);
const char* qsb_last_error(qsb_ctx* ctx);
/* diagnostics of the current cycle: [0] segments that took the full 24-facet geometry path, [1] check-mode (geometry or reaction)
 * disagreements (check mode; must be 0), [2] reaction-table entries scanned, [3] compact geometry enabled,
 * [4] registers per thread, [5] resident blocks per SM, [6] grid size, [7] vault slots used. */
int  qsb_get_diagnostics(qsb_ctx* ctx, uint64_t out[8]);
/* 16 hex digits naming the tracking kernels this library was built from (their sources + tuning knobs): profiler evidence
 * is recorded against it (profiles/dram_traffic.json) and bench.py refuses evidence taken from another kernel */
const char* qsb_kernel_hash(void);
/* number of kernels launched by this context since creation (bench.py's gpu_launches). */
uint64_t qsb_launch_count(qsb_ctx* ctx);

/* ======================================================================================================
 * Drop-in for cycleTracking(MonteCarlo*) (src/main.cc:138-307) on one rank: processing vault -> device,
 * track, census + balance + flux sum -> host model.  Host buffers in, host buffers out.
 * ==================================================================================================== */
int  qsb_mc_cycle_tracking(qsb_mc* mc, qsb_ctx* ctx, qsb_track_stats* stats);
/* the same call in two halves, for drivers that run exchange rounds between ranks in the middle (qsb_track, qsb_send_slab,
 * qsb_put_arrivals ... until no rank sent anything): begin = qsb_cycle_begin + qsb_stream_begin on the host model's vaults,
 * end = qsb_stream_end + balance + flux sum into the host model's tallies. */
int  qsb_mc_tracking_begin(qsb_mc* mc, qsb_ctx* ctx);
int  qsb_mc_tracking_end(qsb_mc* mc, qsb_ctx* ctx);

/* ---- the whole cycle with the population resident on the device ----------------------------------
 *   qsb_mc_cycle_init_resident      cycleInit (src/main.cc:96-121): the host model supplies the global numbers (source
 *                                   weight, per-cell source counts, split factor -- reduced over ranks through the
 *                                   allreduce hook), the device does the per-particle work (qsb_cycle_init_resident);
 *                                   _start/_source/_rr/_split go to the host model's balance.  On the first call the host
 *                                   model's processed vault (if any) is moved to the device.
 *   qsb_mc_cycle_tracking_resident  cycleTracking (src/main.cc:138-307) on one rank: qsb_track + tallies into the host model;
 *                                   the census stays on the device.  Several ranks: qsb_track / exchange rounds by the
 *                                   caller, then qsb_mc_tracking_end_resident.
 *   qsb_mc_cycle_finalize           unchanged (src/main.cc:310-324).
 *   qsb_mc_census_to_host           bring the census back into the host model's processed vault (end of run, census
 *                                   output, or to continue with host-side cycles).
 * Host and device cycles can be mixed freely; only host memory <-> device copies of the whole vault separate them. */
int  qsb_mc_cycle_init_resident(qsb_mc* mc, qsb_ctx* ctx, qsb_cycle_init_result* result /* optional */);
/* The global numbers of the coming cycle's cycleInit, as qsb_mc_cycle_init_resident hands them to the device, without
 * side effects: source_offsets[n_cells+1], source_tally[n_cells] (either may be NULL), the weight of a source particle and
 * the split / roulette factor for a rank whose carried-over census holds n_census particles (reduces over ranks when the
 * deck says loadBalance 0: then every rank calls it). */
int  qsb_mc_source_plan(qsb_mc* mc, int32_t* source_offsets, uint64_t* source_tally, double* source_weight, double* split_factor,
                        uint64_t n_census);
int  qsb_mc_cycle_tracking_resident(qsb_mc* mc, qsb_ctx* ctx, qsb_track_stats* stats /* optional */);
int  qsb_mc_tracking_end_resident(qsb_mc* mc, qsb_ctx* ctx);
int  qsb_mc_census_to_host(qsb_mc* mc, qsb_ctx* ctx);

const char* qsb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QSB_H */
