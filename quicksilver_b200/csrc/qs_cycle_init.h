/* qs_cycle_init.h -- the per-particle arithmetic of cycleInit, shared by the host model (host/MonteCarlo.cc)
 * and the device cycle-init kernel (device/cycle_init_kernels.cu).
 *
 * cycleInit (src/main.cc:96-121) is the step directly in front of the tracking hot path: MC_SourceNow
 * (src/MC_SourceNow.cc:28-133), PopulationControl (src/PopulationControl.cc:20-122) and
 * RouletteLowWeightParticles (src/PopulationControl.cc:127-171).  Every decision in it is taken per particle
 * from that particle's own random-number stream, so it parallelises over particles once the three global
 * numbers (source particle weight, split/roulette factor, weight cut-off) are known.  ONE source for both
 * sides: compiled by g++ (-ffp-contract=off) into the host model, where the golden fixtures dumped from the
 * reference pin it byte for byte, and by nvcc (--fmad=false) into the kernel.
 *
 * Math policy M: QsLibmMath (host only: std::log/sin/cos, the reference's bits) or QsStrictMath
 * (qs_strict_math.h: same bits on host and device).
 */
#ifndef QS_CYCLE_INIT_H
#define QS_CYCLE_INIT_H

#include <stdint.h>

#include "qs_rng.h"
#include "qs_strict_math.h"

#if !defined(__CUDA_ARCH__)
#include <cmath>
#endif

#if defined(__CUDACC__)
#define QS_CI_HD __host__ __device__ __forceinline__
#define QS_CI_MEMBER __host__ __device__ __forceinline__ static
#else
#define QS_CI_HD static inline
#define QS_CI_MEMBER static inline
#endif

#if !defined(__CUDA_ARCH__)
struct QsLibmMath
{
    static inline double log(double x) { return std::log(x); }
    static inline void sincos(double phi, double* s, double* c) { *s = std::sin(phi); *c = std::cos(phi); }
};
#endif

struct QsStrictMath
{
    QS_CI_MEMBER double log(double x) { return qs_strict_log(x); }
    /* qs_strict_sincos reduces arguments in [0, 8); the source's azimuth lies in (-pi, pi): use the symmetry */
    QS_CI_MEMBER void sincos(double phi, double* s, double* c)
    {
        if (phi < 0.0) { qs_strict_sincos(-phi, s, c); *s = -*s; }
        else qs_strict_sincos(phi, s, c);
    }
};

/* facet f of a cell = points {ring[f/4][f%4], ring[f/4][(f%4+1)%4], 8 + f/4} of its 14-point list
 * (src/MC_Domain.cc:41-50 as a closed form; the host model checks it against its table at start-up) */
QS_CI_HD void qs_facet_points(int facet, int* p0, int* p1, int* p2)
{
    const int face = facet >> 2, k = facet & 3;
    /* corner rings of the six faces, one nibble per corner: 1,3,7,5 | 0,4,6,2 | 3,2,6,7 | 0,1,5,4 | 4,5,7,6 | 0,2,3,1 */
    const uint64_t rings_lo = 0x4510762326405731ull;
    const uint32_t rings_hi = 0x13206754u;
    const uint32_t ring = face < 4 ? (uint32_t)(rings_lo >> (16 * face)) & 0xffffu : (rings_hi >> (16 * (face - 4))) & 0xffffu;
    *p0 = (int)((ring >> (4 * k)) & 0xfu);
    *p1 = (int)((ring >> (4 * ((k + 1) & 3))) & 0xfu);
    *p2 = 8 + face;
}

/* 6 x signed volume of the tet (a, b, c, apex)  (src/MCT.cc:627-646) */
QS_CI_HD double qs_tet_det(const double* a, const double* b, const double* c, double ax, double ay, double az)
{
    const double v0x = a[0] - ax, v0y = a[1] - ay, v0z = a[2] - az;
    const double v1x = b[0] - ax, v1y = b[1] - ay, v1z = b[2] - az;
    const double v2x = c[0] - ax, v2y = c[1] - ay, v2z = c[2] - az;
    return v0z * (v1x * v2y - v1y * v2x) + v0y * (v1z * v2x - v1x * v2z) + v0x * (v1y * v2z - v1z * v2y);
}

/* uniform point in a cell: pick one of the 24 centre-apex tets by volume, then fold the unit cube into the tet's
 * barycentric simplex (MCT_Generate_Coordinate_3D_G, src/MCT.cc:143-226).  nodes = the cell's 14 points [14][3]. */
QS_CI_HD void qs_generate_coordinate(uint64_t* seed, const double* nodes, double cell_volume, double out[3])
{
    /* MCT_Cell_Position_3D_G (src/MCT.cc:231-253): mean of the 14 points */
    double cx = 0.0, cy = 0.0, cz = 0.0;
    for (int p = 0; p < 14; ++p) { cx += nodes[3 * p]; cy += nodes[3 * p + 1]; cz += nodes[3 * p + 2]; }
    const double inv = 1.0 / ((double)14);
    cx *= inv; cy *= inv; cz *= inv;

    const double which_volume = qs_rng_sample(seed) * 6.0 * cell_volume;
    double running = 0.0;
    int facet = -1;
    const double *p0 = 0, *p1 = 0, *p2 = 0;
    while (running < which_volume)
    {
        ++facet;
        if (facet == 24) break;
        int i0, i1, i2;
        qs_facet_points(facet, &i0, &i1, &i2);
        p0 = nodes + 3 * i0; p1 = nodes + 3 * i1; p2 = nodes + 3 * i2;
        running += qs_tet_det(p0, p1, p2, cx, cy, cz);
    }
    double r1 = qs_rng_sample(seed), r2 = qs_rng_sample(seed), r3 = qs_rng_sample(seed);
    if (r1 + r2 > 1.0) { r1 = 1.0 - r1; r2 = 1.0 - r2; }
    if (r2 + r3 > 1.0)           { const double t = r3; r3 = 1.0 - r1 - r2; r2 = 1.0 - t; }
    else if (r1 + r2 + r3 > 1.0) { const double t = r3; r3 = r1 + r2 + r3 - 1.0; r1 = 1.0 - r2 - t; }
    const double r4 = 1.0 - r1 - r2 - r3;
    if (!p0) { out[0] = out[1] = out[2] = 0.0; return; }      /* r == 0: the reference bails out with the origin */
    out[0] = (r4 * cx + r1 * p0[0] + r2 * p1[0] + r3 * p2[0]);
    out[1] = (r4 * cy + r1 * p0[1] + r2 * p1[1] + r3 * p2[1]);
    out[2] = (r4 * cz + r1 * p0[2] + r2 * p1[2] + r3 * p2[2]);
}

/* what MC_SourceNow decides for one new particle; everything else of the record is a constant
 * (age, num_segments, num_collisions, breed, species = 0; last_event = census, src/MC_Base_Particle.hh:259) */
typedef struct qs_source_particle
{
    double   coordinate[3];
    double   velocity[3];
    double   kinetic_energy;
    double   num_mean_free_paths;
    double   time_to_census;
    uint64_t random_number_seed;
    uint64_t identifier;
} qs_source_particle;

/* One source particle of a cell (src/MC_SourceNow.cc:86-126).  stream = the cell's running source count (before this
 * particle) + the cell's id (gid << 32, src/MC_Domain.cc:390). */
template <class M>
QS_CI_HD void qs_source_one(uint64_t stream, const double* nodes, double cell_volume, double e_min, double e_max, double dt,
                            qs_source_particle* p)
{
    const double neutron_rest_mass_energy = 9.395656981095e+2;   /* MeV   (src/PhysicalConstants.hh:10-12) */
    const double pi = 3.1415926535897932;
    const double speed_of_light = 2.99792458e+10;                /* cm/s */

    uint64_t s = stream;
    uint64_t seed = qs_rng_spawn(&s);
    p->identifier = s;
    qs_generate_coordinate(&seed, nodes, cell_volume, p->coordinate);

    /* isotropic direction (src/DirectionCosine.cc:5-13) */
    const double gamma = 1.0 - 2.0 * qs_rng_sample(&seed);
    const double sine_gamma = sqrt((1.0 - (gamma * gamma)));
    const double phi = pi * (2.0 * qs_rng_sample(&seed) - 1.0);
    double sin_phi, cos_phi;
    M::sincos(phi, &sin_phi, &cos_phi);
    const double alpha = sine_gamma * cos_phi;
    const double beta = sine_gamma * sin_phi;

    const double e = (e_max - e_min) * qs_rng_sample(&seed) + e_min;
    p->kinetic_energy = e;
    /* src/MC_SourceNow.cc:169-177 */
    const double speed = speed_of_light * sqrt(e * (e + 2.0 * (neutron_rest_mass_energy)) /
                                               ((e + neutron_rest_mass_energy) * (e + neutron_rest_mass_energy)));
    p->velocity[0] = speed * alpha; p->velocity[1] = speed * beta; p->velocity[2] = speed * gamma;
    p->num_mean_free_paths = -1.0 * M::log(qs_rng_sample(&seed));
    p->time_to_census = dt * qs_rng_sample(&seed);
    p->random_number_seed = seed;
}

/* PopulationControlGuts for one particle (src/PopulationControl.cc:66-122).  Draws one number from the particle's
 * stream (the caller skips the call altogether when factor == 1, as the reference does).  Returns -1 if the particle
 * is killed (roulette, factor < 1), else the number of split copies to make (0 when factor < 1); *weight is updated. */
QS_CI_HD int qs_population_control_one(double factor, uint64_t* seed, double* weight)
{
    const double r = qs_rng_sample(seed);
    if (factor < 1)
    {
        if (r > factor) return -1;
        *weight /= factor;
        return 0;
    }
    int copies = (int)floor(factor);
    if (r > (factor - copies)) copies--;
    *weight /= factor;
    return copies;
}

/* RouletteLowWeightParticles for one particle (src/PopulationControl.cc:127-171): returns 0 if it is killed.
 * The caller skips the call when the deck's lowWeightCutoff is not positive. */
QS_CI_HD int qs_roulette_low_weight_one(double cutoff, double weight_cutoff, uint64_t* seed, double* weight)
{
    if (*weight <= weight_cutoff)
    {
        const double r = qs_rng_sample(seed);
        if (r <= cutoff) *weight /= cutoff;
        else return 0;
    }
    return 1;
}

#endif /* QS_CYCLE_INIT_H */
