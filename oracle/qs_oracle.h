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

#ifdef __cplusplus
}
#endif
#endif
