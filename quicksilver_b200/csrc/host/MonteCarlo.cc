// MonteCarlo.cc -- see MonteCarlo.hh.
#include "MonteCarlo.hh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>

#include "../qs_rng.h"
#include "../qs_cycle_init.h"

#include <cuda_runtime_api.h>
#include <cstdlib>
#include <mutex>
#include <unordered_set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace qsb {

// ---- vault storage: page-locked when a CUDA driver is present (see MonteCarlo.hh) ------------------------------
namespace {
std::mutex g_vaultMutex;
std::unordered_set<void*> g_pinnedBlocks;
}

void* vaultAllocate(size_t bytes)
{
    if (bytes == 0) bytes = 1;
    static const bool noPin = std::getenv("QSB_NO_PINNED_VAULTS") != nullptr;
    if (!noPin)
    {
        void* p = nullptr;
        if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess && p)
        {
            std::lock_guard<std::mutex> lock(g_vaultMutex);
            g_pinnedBlocks.insert(p);
            return p;
        }
        cudaGetLastError();      // no driver / no device / out of lockable memory: ordinary memory is fine
    }
    void* p = std::malloc(bytes);
    if (!p) throw std::bad_alloc();
    return p;
}

void vaultFree(void* p)
{
    if (!p) return;
    bool pinned = false;
    {
        std::lock_guard<std::mutex> lock(g_vaultMutex);
        pinned = g_pinnedBlocks.erase(p) != 0;
    }
    if (pinned) cudaFreeHost(p); else std::free(p);
}


MonteCarlo::MonteCarlo(const Parameters& p, int rank_, int nRanks_)
: params(p), rank(rank_), nRanks(nRanks_),
  nuclearData(p.simulationParams.nGroups, p.simulationParams.eMin, p.simulationParams.eMax),
  timeStep(p.simulationParams.dt)
{
    if (nRanks < 1 || rank < 0 || rank >= nRanks) throw std::runtime_error("bad rank / n_ranks");
    if (p.simulationParams.nParticles / (uint64_t)nRanks == 0)
        throw std::runtime_error("not enough particles for each rank");
    // the closed form of the facet -> points table used by the shared cycle-init code must be the mesh builder's table
    for (int f = 0; f < 24; ++f)
    {
        int p0, p1, p2;
        qs_facet_points(f, &p0, &p1, &p2);
        if (p0 != kFacetPoints[f][0] || p1 != kFacetPoints[f][1] || p2 != kFacetPoints[f][2])
            throw std::runtime_error("internal: qs_facet_points disagrees with the facet table");
    }
    initNuclearData(params, nuclearData, materialDatabase);
    initMesh(params, materialDatabase, rank, nRanks, ddc, domain);
    buildImage();
    tallies.censusEnergySpectrum.assign(nuclearData.energies.size(), 0);
    // checkCrossSections (src/initMC.cc:72,392-484): plot data of the cross sections, when a file is named
    if (!params.simulationParams.crossSectionsOut.empty() && rank == 0)
    {
        const std::string name = params.simulationParams.crossSectionsOut + ".dat";
        if (FILE* f = std::fopen(name.c_str(), "w"))
        {
            const std::string text = crossSectionsText(*this);
            std::fwrite(text.data(), 1, text.size(), f);
            std::fclose(f);
        }
    }
}

// Flatten mesh + nuclear data into the arrays of qsb_image (see include/qsb.h for the layout).
void MonteCarlo::buildImage()
{
    const int nDom = (int)domain.size();
    const int nGroups = nuclearData.numGroups;
    const int nMat = (int)materialDatabase.mat.size();
    Flat& f = flat;
    f = Flat();

    f.domainCellOffset.assign(nDom + 1, 0);
    for (int d = 0; d < nDom; ++d)
    {
        f.domainCellOffset[d + 1] = f.domainCellOffset[d] + domain[d].nCells;
        f.domainGid.push_back(domain[d].globalDomain);
    }
    const int nCells = f.domainCellOffset[nDom];
    for (int d = 0; d < nDom; ++d)
    {
        const Domain& D = domain[d];
        f.planes.insert(f.planes.end(), D.planes.begin(), D.planes.end());
        f.nodes.insert(f.nodes.end(), D.nodes.begin(), D.nodes.end());
        f.cellGid.insert(f.cellGid.end(), D.cellGid.begin(), D.cellGid.end());
        f.cellMaterial.insert(f.cellMaterial.end(), D.material.begin(), D.material.end());
        f.cellVolume.insert(f.cellVolume.end(), D.volume.begin(), D.volume.end());
        f.cellId.insert(f.cellId.end(), D.cellId.begin(), D.cellId.end());
        f.faceEvent.insert(f.faceEvent.end(), D.faceEvent.begin(), D.faceEvent.end());
        f.faceAdjDomain.insert(f.faceAdjDomain.end(), D.faceAdjDomain.begin(), D.faceAdjDomain.end());
        f.faceNbrRank.insert(f.faceNbrRank.end(), D.faceNbrRank.begin(), D.faceNbrRank.end());
        for (int c = 0; c < D.nCells; ++c)
            for (int face = 0; face < 6; ++face)
            {
                const size_t k = (size_t)c * 6 + face;
                int adj = D.faceAdjCell[k];
                if (D.faceEvent[k] == QSB_ADJ_TRANSIT_ON)        adj += f.domainCellOffset[D.faceAdjDomain[k]];
                else if (D.faceEvent[k] != QSB_ADJ_TRANSIT_OFF)  adj = f.domainCellOffset[d] + c;   // boundary: itself
                f.faceAdjCell.push_back(adj);
            }
    }

    // nuclear data
    f.energies = nuclearData.energies;
    int maxReact = 1, nIsoTotal = 0;
    for (const Material& m : materialDatabase.mat)
    {
        int n = 0;
        for (int gid : m.isoGid) n += nuclearData.isotopes[gid].nReactions;
        if (n > maxReact) maxReact = n;
        nIsoTotal += (int)m.isoGid.size();
    }
    f.xsTotal.assign((size_t)nMat * nGroups, 0.0);
    f.xsReact.assign((size_t)nMat * nGroups * maxReact, 0.0);
    f.matReactType.assign((size_t)nMat * maxReact, QSB_REACT_UNDEFINED);
    for (int mi = 0; mi < nMat; ++mi)
    {
        const Material& m = materialDatabase.mat[mi];
        const int nIso = (int)m.isoGid.size();
        f.matNIso.push_back(nIso);
        f.matNReact.push_back(nIso ? nuclearData.isotopes[m.isoGid[0]].nReactions : 0);
        f.matMass.push_back(m.mass);
        f.matNuBar.push_back(m.nuBar);
        const double cellNumberDensity = 1.0;                 // src/MC_Domain.cc:387
        bool periodic = true;
        for (int g = 0; g < nGroups; ++g)
        {
            // weightedMacroscopicCrossSection: sum over isotopes of af*density*(sum over reactions of sigma)
            double sum = 0.0;
            int k = 0;
            for (int i = 0; i < nIso; ++i)
            {
                const int gid = m.isoGid[i];
                const double af = m.atomFraction[i];
                if (af == 0.0 || cellNumberDensity == 0.0) sum += 1e-20;
                else sum += af * cellNumberDensity * nuclearData.totalCrossSection(gid, g);
                for (int r = 0; r < nuclearData.isotopes[gid].nReactions; ++r, ++k)
                {
                    const double v = (af == 0.0 || cellNumberDensity == 0.0) ? 1e-20
                                     : af * cellNumberDensity * nuclearData.sigmaOf(gid, r, g);
                    f.xsReact[((size_t)mi * nGroups + g) * maxReact + k] = v;
                    if (g == 0) f.matReactType[(size_t)mi * maxReact + k] = nuclearData.isotopes[gid].reactionType[r];
                    const int nr0 = f.matNReact[mi];
                    if (nuclearData.isotopes[gid].nReactions != nr0 ||
                        std::memcmp(&v, &f.xsReact[((size_t)mi * nGroups + g) * maxReact + r], 8) != 0 ||
                        nuclearData.isotopes[gid].reactionType[r] != nuclearData.isotopes[m.isoGid[0]].reactionType[r])
                        periodic = false;
                }
            }
            f.xsTotal[(size_t)mi * nGroups + g] = sum;
        }
        f.matPeriodic.push_back(periodic ? 1 : 0);
    }

    std::memset(&image, 0, sizeof(image));
    image.abi_version = QSB_ABI_VERSION;
    image.n_domains = nDom; image.n_cells = nCells; image.n_groups = nGroups; image.n_materials = nMat;
    image.n_isotopes = nIsoTotal; image.max_reactions_per_material = maxReact;
    image.my_rank = rank; image.n_ranks = nRanks;
    image.global_nx = params.simulationParams.nx; image.global_ny = params.simulationParams.ny; image.global_nz = params.simulationParams.nz;
    image.global_lx = params.simulationParams.lx; image.global_ly = params.simulationParams.ly; image.global_lz = params.simulationParams.lz;
    image.domain_cell_offset = f.domainCellOffset.data(); image.domain_gid = f.domainGid.data();
    image.planes = f.planes.data(); image.nodes = f.nodes.data(); image.cell_gid = f.cellGid.data();
    image.cell_material = f.cellMaterial.data(); image.cell_volume = f.cellVolume.data(); image.cell_id = f.cellId.data();
    image.face_event = f.faceEvent.data(); image.face_adj_cell = f.faceAdjCell.data();
    image.face_adj_domain = f.faceAdjDomain.data(); image.face_nbr_rank = f.faceNbrRank.data();
    image.energies = f.energies.data(); image.mat_n_isotopes = f.matNIso.data(); image.mat_n_reactions = f.matNReact.data();
    image.mat_mass = f.matMass.data(); image.mat_nu_bar = f.matNuBar.data(); image.mat_react_type = f.matReactType.data();
    image.xs_total = f.xsTotal.data(); image.xs_react = f.xsReact.data(); image.mat_periodic = f.matPeriodic.data();
}

// ---------------------------------------------------------------------------------------------------
// cycleInit
// ---------------------------------------------------------------------------------------------------

void cycleInit(MonteCarlo& mc)
{
    // last cycle's census is this cycle's starting population (src/main.cc:106-110)
    mc.processing.swap(mc.processed);
    mc.processed.clear();
    mc.tallies.balanceTask[QSB_BAL_START] = mc.processing.size();
    mc.tallies.scalarFluxSum = 0.0;
    sourceNow(mc);
    populationControl(mc);
    rouletteLowWeightParticles(mc);
}

namespace {

// Host threads for the cycleInit stages (OpenMP).  Every particle decides from its own stream and lands in a slot fixed by
// prefix sums, so the vault comes out in the same order -- the serial order -- whatever the thread count.
int hostThreads()
{
    static const int n = [] {
        if (const char* e = std::getenv("QSB_HOST_THREADS")) { const int v = std::atoi(e); if (v > 0) return v; }
#ifdef _OPENMP
        return std::min(omp_get_max_threads(), 32);        // memory-bound loops: more threads than that buy nothing
#else
        return 1;
#endif
    }();
    return n;
}
constexpr size_t kParallelThreshold = 1u << 15;        // below this the serial loops are faster than waking a team

template <class M>
void fillSourceParticle(const MonteCarlo& mc, const Domain& d, size_t di, int c, uint64_t stream, double weight, qsb_base_particle& p)
{
    const SimulationParameters& sp = mc.params.simulationParams;
    qs_source_particle sp1;
    qs_source_one<M>(stream, &d.nodes[(size_t)c * 42], d.volume[c], sp.eMin, sp.eMax, mc.timeStep, &sp1);
    std::memset(&p, 0, sizeof(p));
    p.random_number_seed = sp1.random_number_seed;
    p.identifier = sp1.identifier;
    for (int k = 0; k < 3; ++k) { p.coordinate[k] = sp1.coordinate[k]; p.velocity[k] = sp1.velocity[k]; }
    p.kinetic_energy = sp1.kinetic_energy;
    p.domain = (int32_t)di; p.cell = c;
    p.weight = weight;
    p.num_mean_free_paths = sp1.num_mean_free_paths;
    p.time_to_census = sp1.time_to_census;
    p.last_event = QSB_EV_CENSUS;        // MC_Particle's default (src/MC_Base_Particle.hh:259)
    p.species = 0;
}

template <class M>
void sourceCells(MonteCarlo& mc, double weight)
{
    const double dt = mc.timeStep;
    // per-cell counts (src/MC_SourceNow.cc:72-76) and where each cell's particles go in the vault
    struct CellRef { int domain, cell, n; size_t first; };
    std::vector<CellRef> cells;
    size_t total = 0;
    for (size_t di = 0; di < mc.domain.size(); ++di)
    {
        const Domain& d = mc.domain[di];
        for (int c = 0; c < d.nCells; ++c)
        {
            const double cellWeight = d.volume[c] * mc.materialDatabase.mat[d.material[c]].sourceRate * dt;
            const int n = (int)(cellWeight / weight);
            if (n > 0) { cells.push_back(CellRef{ (int)di, c, n, total }); total += (size_t)n; }
        }
    }
    const size_t base = mc.processing.size();
    mc.processing.resize(base + total);
    qsb_base_particle* out = mc.processing.data() + base;
    const long nRef = (long)cells.size();
#pragma omp parallel for schedule(static) num_threads(hostThreads()) if (total >= kParallelThreshold)
    for (long r = 0; r < nRef; ++r)
    {
        const CellRef& ref = cells[r];
        const Domain& d = mc.domain[ref.domain];
        const uint64_t stream0 = d.sourceTally[ref.cell] + d.cellId[ref.cell];
        for (int i = 0; i < ref.n; ++i)
            fillSourceParticle<M>(mc, d, (size_t)ref.domain, ref.cell, stream0 + (uint64_t)i, weight, out[ref.first + i]);
    }
    for (const CellRef& ref : cells) mc.domain[ref.domain].sourceTally[ref.cell] += (uint64_t)ref.n;
    mc.tallies.balanceTask[QSB_BAL_SOURCE] += total;
}

// keep[i] != 0 survivors of `v`, in order, become the vault; the old buffer is kept as scratch for the next time
void compactKept(MonteCarlo& mc, const std::vector<uint8_t>& keep, size_t nKept)
{
    ParticleVault& v = mc.processing;
    const size_t n = v.size();
    ParticleVault& out = mc.scratch;
    out.resize(nKept);
    const int T = std::max(1, std::min<int>(hostThreads(), (int)(n / 4096) + 1));
    std::vector<size_t> first(T + 1, 0);
    const size_t chunk = (n + T - 1) / T;
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t)
    {
        size_t k = 0;
        for (size_t i = t * chunk, e = std::min(n, (t + 1) * chunk); i < e; ++i) k += keep[i] != 0;
        first[t + 1] = k;
    }
    for (int t = 0; t < T; ++t) first[t + 1] += first[t];
#pragma omp parallel for schedule(static, 1) num_threads(T)
    for (int t = 0; t < T; ++t)
    {
        size_t dst = first[t];
        for (size_t i = t * chunk, e = std::min(n, (t + 1) * chunk); i < e; ++i)
            if (keep[i]) out[dst++] = v[i];
    }
    v.swap(out);
    out.clear();
}

} // namespace

// total source weight of the cycle over ALL ranks (src/MC_SourceNow.cc:41-57) -> weight of one source particle (:59-61)
double sourceParticleWeight(MonteCarlo& mc)
{
    const SimulationParameters& sp = mc.params.simulationParams;
    const double dt = mc.timeStep;
    double localWeight = 0;
    for (const Domain& d : mc.domain)
        for (int c = 0; c < d.nCells; ++c)
            localWeight += d.volume[c] * mc.materialDatabase.mat[d.material[c]].sourceRate * dt;
    double totalWeight = localWeight;
    if (!mc.ddc.globalVolumeRate.empty())
    {
        // several ranks: the same sum over ALL cells in global-id order on every rank (no allreduce rounding)
        totalWeight = 0;
        for (double vr : mc.ddc.globalVolumeRate) totalWeight += vr * dt;
    }
    else mc.reduceSum(&totalWeight, 1);

    const double sourceFraction = 0.1;
    return totalWeight / (sourceFraction * sp.nParticles);
}

// split / roulette factor of PopulationControl (src/PopulationControl.cc:20-63) for a rank holding localCount particles
double populationControlFactor(MonteCarlo& mc, uint64_t localCount)
{
    const SimulationParameters& sp = mc.params.simulationParams;
    uint64_t target = sp.nParticles;
    uint64_t globalCount = localCount;
    double factor = 1.0;
    if (sp.loadBalance)
    {
        target = (uint64_t)std::ceil((double)target / (double)mc.nRanks);
        if ((int)localCount != 0) factor = (double)target / (double)(int)localCount;
    }
    else
    {
        mc.reduceSum(&globalCount, 1);
        factor = (double)target / (double)globalCount;
    }
    return factor;
}

void sourceNow(MonteCarlo& mc)
{
    const double weight = sourceParticleWeight(mc);
    mc.sourceParticleWeight = weight;
    if (mc.strictMath) sourceCells<QsStrictMath>(mc, weight);
    else               sourceCells<QsLibmMath>(mc, weight);
}

void populationControl(MonteCarlo& mc)
{
    const uint64_t localCount = mc.processing.size();
    const double factor = populationControlFactor(mc, localCount);
    if (factor == 1.0) return;

    // Every particle decides from its own stream, so the result does not depend on vault order; the
    // reference walks the vault backwards and swap-erases, here survivors keep their order and
    // split copies are appended behind the original population, parent by parent.
    ParticleVault& v = mc.processing;
    Balance& bal = mc.tallies.balanceTask;
    const long n = (long)localCount;
    const bool parallel = localCount >= kParallelThreshold;
    if (factor < 1)
    {
        std::vector<uint8_t> keep(localCount);
        size_t kept = 0;
#pragma omp parallel for schedule(static) reduction(+:kept) num_threads(hostThreads()) if (parallel)
        for (long i = 0; i < n; ++i)
        {
            qsb_base_particle& p = v[i];
            keep[i] = qs_population_control_one(factor, &p.random_number_seed, &p.weight) >= 0;
            kept += keep[i];
        }
        bal[QSB_BAL_RR] += localCount - kept;
        if (kept != localCount) compactKept(mc, keep, kept);
    }
    else
    {
        std::vector<int32_t> copies(localCount);
#pragma omp parallel for schedule(static) num_threads(hostThreads()) if (parallel)
        for (long i = 0; i < n; ++i)
        {
            qsb_base_particle& p = v[i];
            copies[i] = qs_population_control_one(factor, &p.random_number_seed, &p.weight);
        }
        std::vector<size_t> first(localCount + 1, 0);
        for (size_t i = 0; i < localCount; ++i) first[i + 1] = first[i] + (size_t)copies[i];
        const size_t nChildren = first[localCount];
        v.resize(localCount + nChildren);                  // (may move the vault: take the pointer afterwards)
        qsb_base_particle* rec = v.data();
#pragma omp parallel for schedule(dynamic, 4096) num_threads(hostThreads()) if (parallel)
        for (long i = 0; i < n; ++i)
        {
            if (copies[i] <= 0) continue;
            qsb_base_particle& p = rec[i];
            qsb_base_particle child = p;
            for (int k = 0; k < copies[i]; ++k)
            {
                child.random_number_seed = qs_rng_spawn(&p.random_number_seed);
                child.identifier = child.random_number_seed;
                rec[localCount + first[i] + k] = child;
            }
        }
        bal[QSB_BAL_SPLIT] += nChildren;
    }
}

void rouletteLowWeightParticles(MonteCarlo& mc)
{
    const double cutoff = mc.params.simulationParams.lowWeightCutoff;
    if (!(cutoff > 0.0)) return;
    const double weightCutoff = cutoff * mc.sourceParticleWeight;
    ParticleVault& v = mc.processing;
    const long n = (long)v.size();
    std::vector<uint8_t> keep(v.size());
    size_t kept = 0;
#pragma omp parallel for schedule(static) reduction(+:kept) num_threads(hostThreads()) if (v.size() >= kParallelThreshold)
    for (long i = 0; i < n; ++i)
    {
        qsb_base_particle& p = v[i];
        keep[i] = qs_roulette_low_weight_one(cutoff, weightCutoff, &p.random_number_seed, &p.weight) != 0;
        kept += keep[i];
    }
    mc.tallies.balanceTask[QSB_BAL_RR] += v.size() - kept;
    if (kept != v.size()) compactKept(mc, keep, kept);
}

void cycleFinalize(MonteCarlo& mc, Balance& row, double& flux)
{
    // EnergySpectrum::UpdateSpectrum (src/Tallies.cc:97, src/EnergySpectrum.cc:12-35): every particle still held at the end
    // of the cycle -- the census -- counts in its energy group.  A census resident on the device was histogrammed there.
    if (!mc.params.simulationParams.energySpectrum.empty() && !mc.tallies.spectrumDoneThisCycle)
    {
        for (const qsb_base_particle& p : mc.processing) mc.tallies.censusEnergySpectrum[mc.nuclearData.getEnergyGroup(p.kinetic_energy)]++;
        for (const qsb_base_particle& p : mc.processed)  mc.tallies.censusEnergySpectrum[mc.nuclearData.getEnergyGroup(p.kinetic_energy)]++;
    }
    mc.tallies.spectrumDoneThisCycle = false;
    Balance& task = mc.tallies.balanceTask;
    task[QSB_BAL_END] = mc.residentCensus ? mc.residentCensusCount : mc.processed.size();
    row = task;
    mc.reduceSum(row.v, QSB_BAL_COUNT);
    flux = mc.tallies.scalarFluxSum;
    mc.reduceSum(&flux, 1);
    mc.tallies.balanceCumulative.add(row);
    task.reset();
    mc.cycle++;
}

std::vector<uint64_t> globalEnergySpectrum(const MonteCarlo& mc)
{
    std::vector<uint64_t> sum = mc.tallies.censusEnergySpectrum;
    if (!sum.empty()) mc.reduceSum(sum.data(), (int)sum.size());
    return sum;
}

std::string energySpectrumText(const MonteCarlo& mc, const std::vector<uint64_t>& global)
{
    std::string out;
    char line[128];
    for (size_t i = 0; i < mc.nuclearData.energies.size() && i < global.size(); ++i)
    {
        std::snprintf(line, sizeof line, "%d\t%g\t%llu\n", (int)i, mc.nuclearData.energies[i], (unsigned long long)global[i]);
        out += line;
    }
    return out;
}

std::string crossSectionsText(const MonteCarlo& mc)
{
    const NuclearData& nd = mc.nuclearData;
    const int nGroups = (int)nd.energies.size() - 1;
    struct XcData { double absorption = 0., fission = 0., scatter = 0.; };
    std::map<std::string, std::vector<XcData> > table;          // the reference keys a std::map by material name: name order
    for (const Material& m : mc.materialDatabase.mat)
    {
        std::vector<XcData>& xc = table[m.name];
        xc.resize(nGroups);
        const unsigned nIsotopes = (unsigned)m.isoGid.size();
        for (unsigned i = 0; i < nIsotopes; ++i)
        {
            const Isotope& iso = nd.isotopes[m.isoGid[i]];
            for (int r = 0; r < iso.nReactions; ++r)
                for (int g = 0; g < nGroups; ++g)
                {
                    const double v = nd.sigmaOf(m.isoGid[i], r, g) / nIsotopes;
                    switch (iso.reactionType[r])
                    {
                        case QSB_REACT_SCATTER:    xc[g].scatter += v; break;
                        case QSB_REACT_ABSORPTION: xc[g].absorption += v; break;
                        case QSB_REACT_FISSION:    xc[g].fission += v; break;
                        default: break;
                    }
                }
        }
    }
    std::string out = "#group  energy";
    char buf[256];
    for (const auto& kv : table)
    {
        std::snprintf(buf, sizeof buf, "  %s_a  %s_f  %s_s", kv.first.c_str(), kv.first.c_str(), kv.first.c_str());
        out += buf;
    }
    out += "\n";
    for (int g = 0; g < nGroups; ++g)
    {
        std::snprintf(buf, sizeof buf, "%u  %g", (unsigned)g, (nd.energies[g] + nd.energies[g + 1]) / 2.0);
        out += buf;
        for (const auto& kv : table)
        {
            std::snprintf(buf, sizeof buf, "  %g  %g  %g", kv.second[g].absorption, kv.second[g].fission, kv.second[g].scatter);
            out += buf;
        }
        out += "\n";
    }
    return out;
}

} // namespace qsb
