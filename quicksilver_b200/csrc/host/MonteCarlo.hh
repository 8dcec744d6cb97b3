// MonteCarlo.hh -- owner object of the host model (the reference's MonteCarlo god-object,
// src/MonteCarlo.hh:18-47) together with its ParticleVault and Tallies, and the host stages that
// bracket the tracking hot path each cycle: cycleInit (source, population control, roulette) and
// cycleFinalize (balance reduction and bookkeeping).
#ifndef QSB_MONTECARLO_HH
#define QSB_MONTECARLO_HH

#include <cstdint>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/qsb.h"
#include "Mesh.hh"
#include "NuclearData.hh"
#include "Parameters.hh"

namespace qsb {

// Host-side particle vault: one contiguous AoS of MC_Base_Particle-layout records.  The reference's
// fixed-size batches (src/ParticleVaultContainer.hh) exist only to bound kernel launches and carry no
// physics; tallies do not depend on vault order.
//
// Storage is page-locked (cudaHostAlloc) whenever a CUDA driver is present, so that the drop-in call
// qsb_mc_cycle_tracking can stream the vault to the GPU and the census back with plain DMA, overlapped with
// tracking, instead of staging through bounce buffers; on a machine without a GPU it is ordinary heap memory.
// Elements are default-initialised (no zero fill on resize: the census is written by the device).
void* vaultAllocate(size_t bytes);
void  vaultFree(void* p);

template <typename T>
struct VaultAllocator
{
    typedef T value_type;
    VaultAllocator() = default;
    template <typename U> VaultAllocator(const VaultAllocator<U>&) {}
    T* allocate(size_t n) { return static_cast<T*>(vaultAllocate(n * sizeof(T))); }
    void deallocate(T* p, size_t) { vaultFree(p); }
    template <typename U> void construct(U* p) { ::new (static_cast<void*>(p)) U; }
    template <typename U, typename... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
    template <typename U> bool operator==(const VaultAllocator<U>&) const { return true; }
    template <typename U> bool operator!=(const VaultAllocator<U>&) const { return false; }
};
typedef std::vector<qsb_base_particle, VaultAllocator<qsb_base_particle>> ParticleVault;

struct Balance
{
    uint64_t v[QSB_BAL_COUNT] = { 0 };
    uint64_t& operator[](int i) { return v[i]; }
    const uint64_t& operator[](int i) const { return v[i]; }
    void add(const Balance& o) { for (int i = 0; i < QSB_BAL_COUNT; ++i) v[i] += o.v[i]; }
    void reset() { for (int i = 0; i < QSB_BAL_COUNT; ++i) v[i] = 0; }
};

struct Tallies
{
    Balance balanceTask;          // this cycle (replications are summed on the device)
    Balance balanceCumulative;
    double  scalarFluxSum = 0.0;  // this cycle, local
    // EnergySpectrum (src/EnergySpectrum.hh:9-21): per energy-group edge, how many census particles every cycle left in that
    // group, summed over the cycles; only kept when the deck / command line names a spectrum file
    std::vector<uint64_t> censusEnergySpectrum;   // [nGroups+1], local
    bool spectrumDoneThisCycle = false;           // the resident path histograms the census on the device before finalize
};

// MC_Fast_Timer (src/MC_Fast_Timer.hh:27-58): the reference's seven named wall-clock sections, in microseconds.  The library
// times the qsb_mc_* calls that cover a whole section itself; a caller that runs part of a section on its own (the exchange
// rounds of a multi-rank driver) adds its share with qsb_mc_timer_add.
struct FastTimers
{
    enum { Main = 0, CycleInit, CycleTracking, CycleTrackingKernel, CycleTrackingMPI, CycleTrackingTestDone, CycleFinalize, Count };
    double   cumulativeClock[Count] = { 0 }, lastCycleClock[Count] = { 0 };
    uint64_t numCalls[Count] = { 0 };
    double   created = 0.0;            // steady-clock microseconds at construction: start of `main`
    double   trackingStart = -1.0;     // start of the cycle's tracking section when it spans several calls
    void add(int t, double microseconds, uint64_t calls) { cumulativeClock[t] += microseconds; lastCycleClock[t] += microseconds; numCalls[t] += calls; }
    void clearLastCycle() { for (double& v : lastCycleClock) v = 0.0; }
};

class MonteCarlo
{
public:
    MonteCarlo(const Parameters& params, int rank, int nRanks);

    Parameters        params;
    int               rank, nRanks;
    NuclearData       nuclearData;
    MaterialDatabase  materialDatabase;
    DecompositionInfo ddc;
    std::vector<Domain> domain;
    Tallies           tallies;
    FastTimers        timers;
    ParticleVault     processing, processed;
    ParticleVault     scratch;          // destination of cycleInit's compactions; swapped with `processing`, buffer kept
    double            timeStep;
    int               cycle = 0;
    double            sourceParticleWeight = 0.0;
    // 0: libm log/sin/cos in MC_SourceNow (the reference's bits); 1: the portable functions of qs_strict_math.h, which the
    // device cycle-init kernel evaluates to the same bits (qsb_mc_set_strict_math)
    bool              strictMath = false;

    // ---- device-resident cycles (qsb_mc_cycle_init_resident / qsb_mc_cycle_tracking_resident, capi_host.cc) ----
    // The particle population lives in the device context's vaults from one cycle to the next; the host model keeps the
    // counts it needs for the balance bookkeeping and the source plan it hands to the device.
    bool                  residentCensus = false;   // the census of the last cycle is in the device's census vault
    uint64_t              residentCensusCount = 0;
    std::vector<int32_t>  sourceOffsets;            // [nCells+1] prefix sum of the per-cell source counts (flat cell order)
    std::vector<uint64_t> sourceTallyFlat;          // [nCells] scratch: per-cell running source counts handed to the device
    double                sourcePlanWeight = -1.0;  // source particle weight the plan was built for
    double                cachedSourceWeight = 0.0, cachedSourceWeightDt = -1.0;
    uint64_t              devicePlanId = ~0ull;      // plan id the device context `devicePlanCtx` holds
    const void*           devicePlanCtx = nullptr;
    uint64_t              sourcePlanId = 0;         // bumped whenever the plan or the host-side tallies change under the device

    qsb_allreduce_fn  allreduce = nullptr;
    void*             allreduceUser = nullptr;
    void reduceSum(double* v, int n) const { if (allreduce && nRanks > 1) allreduce(allreduceUser, v, n, 0); }
    void reduceSum(uint64_t* v, int n) const { if (allreduce && nRanks > 1) allreduce(allreduceUser, v, n, 1); }

    // the flattened image handed to the device context / oracle; arrays live in `flat`
    void buildImage();
    qsb_image image;
    struct Flat
    {
        std::vector<int32_t> domainCellOffset, domainGid, cellGid, cellMaterial, faceAdjCell, faceAdjDomain, faceNbrRank;
        std::vector<double>  planes, nodes, cellVolume, energies, matMass, matNuBar, xsTotal, xsReact;
        std::vector<uint64_t> cellId;
        std::vector<uint8_t> faceEvent, matReactType, matPeriodic;
        std::vector<int32_t> matNIso, matNReact;
    } flat;

    std::string lastError;
};

// src/main.cc:96-121
void cycleInit(MonteCarlo& mc);
// src/MC_SourceNow.cc:28-133
void sourceNow(MonteCarlo& mc);
// the two global numbers of cycleInit: weight of one source particle (src/MC_SourceNow.cc:41-61) and the split / roulette
// factor for a rank holding localCount particles (src/PopulationControl.cc:20-63); both reduce over ranks
double sourceParticleWeight(MonteCarlo& mc);
double populationControlFactor(MonteCarlo& mc, uint64_t localCount);
// src/PopulationControl.cc:20-122, :127-171
void populationControl(MonteCarlo& mc);
void rouletteLowWeightParticles(MonteCarlo& mc);
// src/main.cc:310-324 + src/Tallies.cc:25-98.  Returns this cycle's global balance and flux sum.
void cycleFinalize(MonteCarlo& mc, Balance& globalRow, double& globalFlux);
// EnergySpectrum::PrintSpectrum (src/EnergySpectrum.cc:37-62): global counts; the text of <energySpectrum>.dat
std::vector<uint64_t> globalEnergySpectrum(const MonteCarlo& mc);
std::string energySpectrumText(const MonteCarlo& mc, const std::vector<uint64_t>& global);
// checkCrossSections (src/initMC.cc:392-484): the text of <crossSectionsOut>.dat
std::string crossSectionsText(const MonteCarlo& mc);

} // namespace qsb
#endif
