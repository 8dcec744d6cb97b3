/* qs_oracle.h -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's cycle-tracking hot
 * path (see qs_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this; the product (quicksilver_b200/) never does. */
#ifndef QS_ORACLE_H
#define QS_ORACLE_H

#include <stdint.h>
#include "../include/qsb.h"

#ifdef __cplusplus
extern "C" {
#endif

/* a particle in transit between ranks: the reference ships MC_Base_Particle only
 * (src/MC_Base_Particle.cc:60-85); the direction cosine rides along so that an N-domain run reproduces
 * the 1-domain history bit for bit. direction_cosine[0] = NaN means "derive from velocity". */
typedef struct qso_exchange_record {
    qsb_base_particle p;
    double direction_cosine[3];
} qso_exchange_record;

typedef struct qso_io {
    /* in */
    const qsb_base_particle*   initial;     uint64_t n_initial;    /* processing vault                      */
    const qso_exchange_record* arrivals;    uint64_t n_arrivals;   /* particles received from other ranks   */
    /* out (caller allocated) */
    qsb_base_particle*   census;   uint64_t census_cap;  uint64_t n_census;
    qso_exchange_record* sends;    int32_t* send_rank;   uint64_t send_cap;  uint64_t n_sends;
    uint64_t balance[QSB_BAL_COUNT];        /* accumulated into                                           */
    double*  flux;                          /* [n_cells][n_groups], accumulated into; may be NULL          */
    uint64_t n_processed;                   /* queue entries consumed, secondaries included                */
    /* diagnostics */
    uint64_t n_retry_moves, n_forced_collisions, n_reaction_lookups;
} qso_io;

/* Track every particle to census / absorption / escape / rank exit, secondaries included.
 * strict_math != 0 uses quicksilver_b200/csrc/qs_strict_math.h for log/sin/cos (bit-comparable with the
 * device validation build); 0 uses libm (bit-comparable with the reference binary on this host).
 * n_threads <= 1 runs serially (deterministic flux summation order). Returns 0, or <0 on overflow. */
int qso_track(const qsb_image* image, double time_step, int strict_math, int n_threads, qso_io* io);

int qso_energy_group(const qsb_image* image, double energy);

/* cycleInit (src/main.cc:96-121), the step in front of the path, restated stage by stage in the reference's own order:
 * last cycle's census becomes the vault, MC_SourceNow appends this cycle's source particles (src/MC_SourceNow.cc:74-126),
 * PopulationControlGuts marches backwards over the vault, erase-swapping the killed and appending split copies
 * (src/PopulationControl.cc:66-122), RouletteLowWeightParticles does the same for low weights (:127-171).
 * The caller supplies the numbers that need the deck or other ranks: the per-cell source counts as a prefix sum
 * (src/MC_SourceNow.cc:72-76), the cells' running source counts, the source particle weight (:59-61) and the split factor
 * (src/PopulationControl.cc:32-57; 1.0 = stage skipped, :60).  Returns 0, or -1 if `out` is too small. */
typedef struct qso_cycle_init_io {
    /* in */
    const qsb_base_particle* census;  uint64_t n_census;
    const int32_t*  source_offsets;       /* [n_cells+1] */
    const uint64_t* source_tally;         /* [n_cells]   */
    double source_weight, e_min, e_max, time_step, split_factor, low_weight_cutoff;
    /* out (caller allocated) */
    qsb_base_particle* out;  uint64_t out_cap;  uint64_t n_out;
    uint64_t n_source, n_rr, n_split;
} qso_cycle_init_io;
int qso_cycle_init(const qsb_image* image, int strict_math, qso_cycle_init_io* io);

#ifdef __cplusplus
}
#endif
#endif
