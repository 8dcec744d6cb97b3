// capi_host.cc -- extern "C" entry points of the host model (qsb_mc_*), see include/qsb.h.
// Every entry point catches exceptions and reports through the return code + qsb_mc_last_error.
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>
#include <climits>

#include "MonteCarlo.hh"

using namespace qsb;

struct qsb_mc
{
    MonteCarlo* mc = nullptr;
    std::string error;
    ParticleVault scratch;
};

namespace {
thread_local std::string g_createError;

double nowMicroseconds()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// The source plan of a cycle: weight of one source particle (src/MC_SourceNow.cc:41-61) and the per-cell counts
// (int)(cellWeight / weight) (:72-76) as a prefix sum over flat cells.  Both are functions of the deck and the time step
// only, so they are evaluated once (the weight walks every cell of every rank) and rebuilt when the time step changes.
double buildSourcePlan(MonteCarlo& mc)
{
    if (mc.cachedSourceWeightDt != mc.timeStep) { mc.cachedSourceWeight = sourceParticleWeight(mc); mc.cachedSourceWeightDt = mc.timeStep; }
    const double weight = mc.cachedSourceWeight;
    const size_t nCells = (size_t)mc.image.n_cells;
    if (mc.sourceOffsets.size() != nCells + 1 || weight != mc.sourcePlanWeight)
    {
        mc.sourceOffsets.assign(nCells + 1, 0);
        size_t flat = 0;
        for (const Domain& d : mc.domain)
            for (int c = 0; c < d.nCells; ++c, ++flat)
            {
                const double cellWeight = d.volume[c] * mc.materialDatabase.mat[d.material[c]].sourceRate * mc.timeStep;
                const int n = (int)(cellWeight / weight);
                const int64_t next = (int64_t)mc.sourceOffsets[flat] + (n > 0 ? n : 0);
                if (next > INT32_MAX) throw std::runtime_error("more than 2^31 source particles on one rank");
                mc.sourceOffsets[flat + 1] = (int32_t)next;
            }
        mc.sourcePlanWeight = weight;
        mc.sourcePlanId++;
    }
    return weight;
}

// times one section of the reference's timer table for the life of the object
struct SectionTimer
{
    FastTimers& t; int which; double t0;
    SectionTimer(FastTimers& timers, int w) : t(timers), which(w), t0(nowMicroseconds()) {}
    ~SectionTimer() { t.add(which, nowMicroseconds() - t0, 1); }
};

template <typename F>
int guarded(qsb_mc* h, F&& body)
{
    if (!h || !h->mc) return QSB_ERR_ARG;
    try { return body(*h->mc); }
    catch (const std::bad_alloc&) { h->error = "out of host memory"; return QSB_ERR_INTERNAL; }
    catch (const std::exception& e) { h->error = e.what(); return QSB_ERR_INTERNAL; }
    catch (...) { h->error = "unknown error"; return QSB_ERR_INTERNAL; }
}
}

extern "C" {

const char* qsb_version(void) { return "quicksilver_b200 0.1 (abi 1)"; }

int qsb_mc_create(int argc, const char* const* argv, int rank, int n_ranks, qsb_mc** out)
{
    if (!out || argc < 0 || (argc > 0 && !argv)) return QSB_ERR_ARG;
    *out = nullptr;
    qsb_mc* h = new (std::nothrow) qsb_mc;
    if (!h) return QSB_ERR_INTERNAL;
    try
    {
        Parameters params = getParameters(argc, argv);
        h->mc = new MonteCarlo(params, rank, n_ranks);
        h->mc->timers.created = nowMicroseconds();            // MC_FASTTIMER_START(main) "once mcco exists" (src/main.cc:50)
    }
    catch (const std::exception& e)
    {
        g_createError = e.what();
        delete h;
        return QSB_ERR_INPUT;
    }
    *out = h;
    return QSB_OK;
}

int qsb_mc_destroy(qsb_mc* h)
{
    if (!h) return QSB_ERR_ARG;
    delete h->mc;
    delete h;
    return QSB_OK;
}

const char* qsb_mc_last_error(qsb_mc* h) { return h ? h->error.c_str() : g_createError.c_str(); }

int qsb_mc_set_allreduce(qsb_mc* h, qsb_allreduce_fn fn, void* user)
{
    return guarded(h, [&](MonteCarlo& mc) { mc.allreduce = fn; mc.allreduceUser = user; return QSB_OK; });
}

int qsb_mc_print_parameters(qsb_mc* h, char* buf, uint64_t cap, uint64_t* needed)
{
    return guarded(h, [&](MonteCarlo& mc) {
        const std::string s = printParameters(mc.params);
        if (needed) *needed = s.size() + 1;
        if (buf && cap) { std::snprintf(buf, cap, "%s", s.c_str()); }
        return QSB_OK;
    });
}

int qsb_mc_get_image(qsb_mc* h, qsb_image* out)
{
    if (!out) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) { *out = mc.image; return QSB_OK; });
}

int qsb_mc_get_int(qsb_mc* h, const char* key, int64_t* out)
{
    if (!key || !out) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) {
        const SimulationParameters& s = mc.params.simulationParams;
        const std::string k = key;
        if      (k == "nSteps") *out = s.nSteps;
        else if (k == "nParticles") *out = (int64_t)s.nParticles;
        else if (k == "nx") *out = s.nx; else if (k == "ny") *out = s.ny; else if (k == "nz") *out = s.nz;
        else if (k == "xDom") *out = s.xDom; else if (k == "yDom") *out = s.yDom; else if (k == "zDom") *out = s.zDom;
        else if (k == "nGroups") *out = s.nGroups;
        else if (k == "loadBalance") *out = s.loadBalance;
        else if (k == "coralBenchmark") *out = s.coralBenchmark;
        else if (k == "cycleTimers") *out = s.cycleTimers;
        else if (k == "nBatches") *out = (int64_t)s.nBatches; else if (k == "batchSize") *out = (int64_t)s.batchSize;
        else if (k == "bTally") *out = s.balanceTallyReplications;
        else if (k == "fTally") *out = s.fluxTallyReplications;
        else if (k == "cycle") *out = mc.cycle;
        else if (k == "nDomains") *out = (int64_t)mc.domain.size();
        else if (k == "nCells") *out = mc.image.n_cells;
        else if (k == "nProcessing") *out = (int64_t)mc.processing.size();
        else if (k == "nProcessed") *out = (int64_t)(mc.residentCensus ? mc.residentCensusCount : mc.processed.size());
        else if (k == "residentCensus") *out = mc.residentCensus ? 1 : 0;
        else { mc.lastError = "unknown integer key " + k; return (int)QSB_ERR_ARG; }
        return (int)QSB_OK;
    });
}

int qsb_mc_get_double(qsb_mc* h, const char* key, double* out)
{
    if (!key || !out) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) {
        const SimulationParameters& s = mc.params.simulationParams;
        const std::string k = key;
        if      (k == "dt") *out = s.dt;
        else if (k == "lx") *out = s.lx; else if (k == "ly") *out = s.ly; else if (k == "lz") *out = s.lz;
        else if (k == "eMin") *out = s.eMin; else if (k == "eMax") *out = s.eMax;
        else if (k == "lowWeightCutoff") *out = s.lowWeightCutoff;
        else if (k == "source_particle_weight") *out = mc.sourceParticleWeight;
        else if (k == "max_nu_bar")
        {
            double m = 0; for (const Material& mat : mc.materialDatabase.mat) if (mat.nuBar > m) m = mat.nuBar;
            *out = m;
        }
        else return (int)QSB_ERR_ARG;
        return (int)QSB_OK;
    });
}

int qsb_mc_cycle_init(qsb_mc* h)
{
    return guarded(h, [&](MonteCarlo& mc) {
        if (mc.residentCensus)
        {
            h->error = "qsb_mc_cycle_init: the census of the last cycle is resident on the device; call qsb_mc_census_to_host first "
                       "(or continue with qsb_mc_cycle_init_resident)";
            return (int)QSB_ERR_STATE;
        }
        mc.timers.clearLastCycle();                           // Last_Cycle_Report clears them at the end of a cycle (src/MC_Fast_Timer.cc:151)
        SectionTimer timer(mc.timers, FastTimers::CycleInit);
        cycleInit(mc);
        mc.sourcePlanId++;          // the host advanced the cells' source counts: a device copy of them is stale
        return (int)QSB_OK;
    });
}

int qsb_mc_source_plan(qsb_mc* h, int32_t* source_offsets, uint64_t* source_tally, double* source_weight, double* split_factor,
                       uint64_t n_census)
{
    return guarded(h, [&](MonteCarlo& mc) {
        const double weight = buildSourcePlan(mc);
        const size_t nCells = (size_t)mc.image.n_cells;
        if (source_offsets) std::memcpy(source_offsets, mc.sourceOffsets.data(), (nCells + 1) * sizeof(int32_t));
        if (source_tally)
        {
            size_t flat = 0;
            for (const Domain& d : mc.domain)
                for (int c = 0; c < d.nCells; ++c, ++flat) source_tally[flat] = d.sourceTally[c];
        }
        if (source_weight) *source_weight = weight;
        if (split_factor) *split_factor = populationControlFactor(mc, n_census + (uint64_t)mc.sourceOffsets[nCells]);
        return QSB_OK;
    });
}

int qsb_mc_set_strict_math(qsb_mc* h, int on)
{
    return guarded(h, [&](MonteCarlo& mc) { mc.strictMath = on != 0; return QSB_OK; });
}

int qsb_mc_processing(qsb_mc* h, const qsb_base_particle** aos, uint64_t* n)
{
    if (!aos || !n) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) { *aos = mc.processing.data(); *n = mc.processing.size(); return QSB_OK; });
}

int qsb_mc_processed(qsb_mc* h, const qsb_base_particle** aos, uint64_t* n)
{
    if (!aos) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) { *aos = mc.processed.data(); if (n) *n = mc.processed.size(); return QSB_OK; });
}

int qsb_mc_set_tracking_result(qsb_mc* h, const qsb_base_particle* census, uint64_t n_census,
                               const uint64_t balance[QSB_BAL_COUNT], double scalar_flux_sum)
{
    if ((n_census && !census) || !balance) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) {
        mc.processed.assign(census, census + n_census);
        mc.processing.clear();
        mc.residentCensus = false;
        static const int tracked[] = { QSB_BAL_ABSORB, QSB_BAL_CENSUS, QSB_BAL_ESCAPE, QSB_BAL_COLLISION, QSB_BAL_FISSION,
                                       QSB_BAL_PRODUCE, QSB_BAL_SCATTER, QSB_BAL_NUM_SEGMENTS };
        for (int i : tracked) mc.tallies.balanceTask[i] += balance[i];
        mc.tallies.scalarFluxSum += scalar_flux_sum;
        return QSB_OK;
    });
}

int qsb_mc_cycle_finalize(qsb_mc* h, uint64_t row[QSB_BAL_COUNT], double* flux)
{
    return guarded(h, [&](MonteCarlo& mc) {
        SectionTimer timer(mc.timers, FastTimers::CycleFinalize);
        Balance r; double f = 0;
        cycleFinalize(mc, r, f);
        if (row) std::memcpy(row, r.v, sizeof(r.v));
        if (flux) *flux = f;
        return QSB_OK;
    });
}

int qsb_mc_energy_spectrum(qsb_mc* h, uint64_t* counts, uint64_t cap, uint64_t* n)
{
    return guarded(h, [&](MonteCarlo& mc) {
        const std::vector<uint64_t> g = globalEnergySpectrum(mc);
        if (n) *n = g.size();
        if (counts) { if (cap < g.size()) { h->error = "spectrum buffer too small"; return (int)QSB_ERR_CAPACITY; } std::memcpy(counts, g.data(), g.size() * sizeof(uint64_t)); }
        return (int)QSB_OK;
    });
}

int qsb_mc_write_energy_spectrum(qsb_mc* h)
{
    return guarded(h, [&](MonteCarlo& mc) {
        const std::string& name = mc.params.simulationParams.energySpectrum;
        if (name.empty()) return (int)QSB_OK;                         // src/EnergySpectrum.cc:39
        const std::vector<uint64_t> g = globalEnergySpectrum(mc);    // every rank reduces, rank 0 writes
        if (mc.rank != 0) return (int)QSB_OK;
        const std::string file = name + ".dat";
        FILE* f = std::fopen(file.c_str(), "w");
        if (!f) { h->error = "cannot write " + file; return (int)QSB_ERR_INPUT; }
        const std::string text = energySpectrumText(mc, g);
        std::fwrite(text.data(), 1, text.size(), f);
        std::fclose(f);
        return (int)QSB_OK;
    });
}

int qsb_mc_cross_sections_text(qsb_mc* h, char* buf, uint64_t cap, uint64_t* needed)
{
    return guarded(h, [&](MonteCarlo& mc) {
        const std::string s = crossSectionsText(mc);
        if (needed) *needed = s.size() + 1;
        if (buf && cap) std::snprintf(buf, cap, "%s", s.c_str());
        return (int)QSB_OK;
    });
}

int qsb_mc_cumulative_balance(qsb_mc* h, uint64_t out[QSB_BAL_COUNT])
{
    if (!out) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) { std::memcpy(out, mc.tallies.balanceCumulative.v, sizeof(uint64_t) * QSB_BAL_COUNT); return QSB_OK; });
}

// src/CoralBenchmark.cc:17-226, statement by statement; the text is the reference's.
int qsb_mc_coral_benchmark_report(qsb_mc* h, const double* fluence, uint64_t n_cells, char* buf, uint64_t cap, uint64_t* needed,
                                  int32_t* passed)
{
    return guarded(h, [&](MonteCarlo& mc) {
        std::string out;
        int n_pass = 0;
        char tmp[512];
        const int which = mc.params.simulationParams.coralBenchmark;
        if (which)
        {
            const Balance& b = mc.tallies.balanceCumulative;
            const bool report = mc.rank == 0;
            {   // BalanceRatioTest (:46-115)
                const uint64_t absorb = b[QSB_BAL_ABSORB], fission = b[QSB_BAL_FISSION], scatter = b[QSB_BAL_SCATTER];
                double absorbRatio = 0.04, fissionRatio = 0.05, scatterRatio = 1, percent_tolerance = 1.0;
                if (which == 2) { fissionRatio = 0.075; scatterRatio = 0.830; absorbRatio = 0.094; percent_tolerance = 1.1; }
                const double tolerance = percent_tolerance / 100.0;
                const double a2s = std::abs((absorb / absorbRatio) * (scatterRatio / scatter) - 1);
                const double a2f = std::abs((absorb / absorbRatio) * (fissionRatio / fission) - 1);
                const double s2a = std::abs((scatter / scatterRatio) * (absorbRatio / absorb) - 1);
                const double s2f = std::abs((scatter / scatterRatio) * (fissionRatio / fission) - 1);
                const double f2a = std::abs((fission / fissionRatio) * (absorbRatio / absorb) - 1);
                const double f2s = std::abs((fission / fissionRatio) * (scatterRatio / scatter) - 1);
                const bool pass = !(a2s > tolerance || a2f > tolerance || s2a > tolerance || s2f > tolerance || f2a > tolerance || f2s > tolerance);
                out += "\nTesting Ratios for Absorbtion, Fission, and Scattering are maintained\n";
                if (pass)
                {
                    std::snprintf(tmp, sizeof tmp, "PASS:: Absorption / Fission / Scatter Ratios maintained with %g%% tolerance\n", tolerance * 100.0);
                    out += tmp;
                }
                else
                {
                    std::snprintf(tmp, sizeof tmp, "FAIL:: Absorption / Fission / Scatter Ratios NOT maintained with %g%% tolerance\n", tolerance * 100.0);
                    out += tmp;
                    std::snprintf(tmp, sizeof tmp, "absorb:  %12llu\t%g\nscatter: %12llu\t%g\nfission: %12llu\t%g\n", (unsigned long long)absorb, absorbRatio,
                                  (unsigned long long)scatter, scatterRatio, (unsigned long long)fission, fissionRatio);
                    out += tmp;
                    const char* names[6] = { "Absorb to Scatter: ", "Absorb to Fission: ", "Scatter to Absorb: ", "Scatter to Fission:", "Fission to Absorb: ", "Fission to Scatter:" };
                    const double vals[6] = { a2s, a2f, s2a, s2f, f2a, f2s };
                    for (int i = 0; i < 6; ++i) { std::snprintf(tmp, sizeof tmp, "Relative %s %g < %g\n", names[i], vals[i], tolerance); out += tmp; }
                }
                n_pass += pass;
            }
            {   // BalanceEventTest (:117-147)
                const uint64_t facetCrossing = b[QSB_BAL_NUM_SEGMENTS] - b[QSB_BAL_CENSUS] - b[QSB_BAL_COLLISION];
                const double ratio = std::abs((double(facetCrossing) / double(b[QSB_BAL_COLLISION])) - 1);
                const double tolerance = 1.0;
                const bool pass = !(ratio > (tolerance / 100.0));
                out += "\nTesting balance between number of facet crossings and reactions\n";
                if (pass) std::snprintf(tmp, sizeof tmp, "PASS:: Collision to Facet Crossing Ratio maintained even balanced within %g%% tolerance\n", tolerance);
                else std::snprintf(tmp, sizeof tmp, " FAIL:: Collision to Facet Crossing Ratio balanced NOT maintained within %g%% tolerance\n"
                                   "\tFacet Crossing: %llu\tCollision: %llu\tRatio: %g\n", tolerance, (unsigned long long)facetCrossing,
                                   (unsigned long long)b[QSB_BAL_COLLISION], ratio);
                out += tmp;
                n_pass += pass;
            }
            {   // MissingParticleTest (:149-171)
                const uint64_t gains = b[QSB_BAL_START] + b[QSB_BAL_SOURCE] + b[QSB_BAL_PRODUCE] + b[QSB_BAL_SPLIT];
                const uint64_t losses = b[QSB_BAL_ABSORB] + b[QSB_BAL_CENSUS] + b[QSB_BAL_ESCAPE] + b[QSB_BAL_RR] + b[QSB_BAL_FISSION];
                out += "\nTest for lost / unaccounted for particles in this simulation\n";
                out += gains == losses ? "PASS:: No Particles Lost During Run\n" : "FAIL:: Particles Were Lost During Run, test for done should have failed\n";
                n_pass += gains == losses;
            }
            {   // FluenceTest (:174-226): this rank's cells (one Fluence domain per mesh domain), max over ranks
                out += "\nTest Fluence for homogeneity across cells\n";
                double max_diff = 0.0;
                const int nDom = (int)mc.flat.domainCellOffset.size() - 1;
                for (int d = 0; d < nDom && fluence; ++d)
                {
                    const size_t first = (size_t)mc.flat.domainCellOffset[d], last = (size_t)mc.flat.domainCellOffset[d + 1];
                    if (last > n_cells) break;
                    double local_sum = 0.0;
                    for (size_t c = first; c < last; ++c) local_sum += fluence[c];
                    const double average = local_sum / (double)(int)(last - first);
                    for (size_t c = first; c < last; ++c)
                    {
                        const double v = fluence[c];
                        const double percent_diff = (((v > average) ? v - average : average - v) / ((v + average) / 2.0)) * 100;
                        max_diff = (max_diff > percent_diff) ? max_diff : percent_diff;
                    }
                }
                if (mc.allreduce && mc.nRanks > 1) mc.allreduce(mc.allreduceUser, &max_diff, 1, 2);
                const double percent_tolerance = 6.0;
                if (max_diff > percent_tolerance)
                {
                    std::snprintf(tmp, sizeof tmp, "FAIL:: Fluence not homogenous across cells within %g%% tolerance\n"
                                  "\tTry running more particles or more cycles to see if Max Percent Difference goes down.\n"
                                  "\tCurrent Max Percent Diff: %4.1f%%\n", percent_tolerance, max_diff);
                }
                else
                {
                    std::snprintf(tmp, sizeof tmp, "PASS:: Fluence is homogenous across cells with %g%% tolerance\n", percent_tolerance);
                    n_pass += 1;
                }
                out += tmp;
            }
            if (!report) out.clear();
        }
        if (passed) *passed = n_pass;
        if (needed) *needed = out.size() + 1;
        if (buf && cap) std::snprintf(buf, cap, "%s", out.c_str());
        return QSB_OK;
    });
}

int qsb_mc_timer_add(qsb_mc* h, int timer, double microseconds, uint64_t calls)
{
    if (timer < 0 || timer >= FastTimers::Count) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) { mc.timers.add(timer, microseconds, calls); return QSB_OK; });
}

int qsb_mc_get_timer(qsb_mc* h, int timer, double* cumulative_microseconds, uint64_t* calls)
{
    if (timer < 0 || timer >= FastTimers::Count) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) {
        if (cumulative_microseconds) *cumulative_microseconds = mc.timers.cumulativeClock[timer];
        if (calls) *calls = mc.timers.numCalls[timer];
        return QSB_OK;
    });
}

// MC_Fast_Timer_Container::Cumulative_Report / Last_Cycle_Report (src/MC_Fast_Timer.cc:58-152), the reference's wording and
// column formats.  min / avg / max / stddev are over ranks (allreduce hook; single rank: all equal, stddev 0).
int qsb_mc_format_timer_report(qsb_mc* h, int last_cycle, char* buf, uint64_t cap, uint64_t* needed)
{
    return guarded(h, [&](MonteCarlo& mc) {
        static const char* names[FastTimers::Count] = { "main", "cycleInit", "cycleTracking", "cycleTracking_Kernel", "cycleTracking_MPI",
                                                        "cycleTracking_Test_Done", "cycleFinalize" };
        FastTimers& t = mc.timers;
        double clock[FastTimers::Count];
        for (int i = 0; i < FastTimers::Count; ++i) clock[i] = std::floor(last_cycle ? t.lastCycleClock[i] : t.cumulativeClock[i]);   // the reference counts whole microseconds
        uint64_t calls[FastTimers::Count];
        for (int i = 0; i < FastTimers::Count; ++i) calls[i] = t.numCalls[i];
        if (!last_cycle) { clock[FastTimers::Main] = std::floor(nowMicroseconds() - t.created); calls[FastTimers::Main] = 1; }  // MC_FASTTIMER_STOP(main)
        double sum[FastTimers::Count], sumsq[FastTimers::Count], mx[FastTimers::Count], negmin[FastTimers::Count];
        for (int i = 0; i < FastTimers::Count; ++i) { sum[i] = clock[i]; sumsq[i] = clock[i] * clock[i]; mx[i] = clock[i]; negmin[i] = -clock[i]; }
        if (mc.allreduce && mc.nRanks > 1)
        {
            mc.allreduce(mc.allreduceUser, sum, FastTimers::Count, 0);
            mc.allreduce(mc.allreduceUser, sumsq, FastTimers::Count, 0);
            mc.allreduce(mc.allreduceUser, mx, FastTimers::Count, 2);
            mc.allreduce(mc.allreduceUser, negmin, FastTimers::Count, 2);
        }
        std::string out;
        char line[256];
        const char* what = last_cycle ? "Last Cycle" : "Cumulative";
        std::snprintf(line, sizeof line, "\n%-25s %12s %12s %12s %12s %12s %12s\n", "Timer", what, what, what, what, what, what); out += line;
        std::snprintf(line, sizeof line, "%-25s %12s %12s %12s %12s %12s %12s\n", "Name", "number", "microSecs", "microSecs", "microSecs", "microSecs", "Efficiency"); out += line;
        std::snprintf(line, sizeof line, "%-25s %12s %12s %12s %12s %12s %12s\n", "", "of calls", "min", "avg", "max", "stddev", "Rating"); out += line;
        for (int i = 0; i < FastTimers::Count; ++i)
        {
            const double ave = std::floor(sum[i] / mc.nRanks);                       // integer average, as the reference's uint64 division
            const double var = sumsq[i] / mc.nRanks - (sum[i] / mc.nRanks) * (sum[i] / mc.nRanks);
            std::snprintf(line, sizeof line, "%-25s %12lu %12.3e %12.3e %12.3e %12.3e %12.2f\n", names[i], (unsigned long)calls[i], -negmin[i], ave, mx[i],
                          std::sqrt(var > 0.0 ? var : 0.0), (100.0 * ave) / (mx[i] + 1.0e-80));
            out += line;
        }
        if (!last_cycle)
        {
            const double segs = (double)mc.tallies.balanceCumulative[QSB_BAL_NUM_SEGMENTS];
            std::snprintf(line, sizeof line, "%-25s %12.3e %-25s\n", "Figure Of Merit", segs / (mx[FastTimers::CycleTracking] * 1e-6),
                          "[Num Segments / Cycle Tracking Time]");
            out += line;
        }
        if (mc.rank != 0) out.clear();
        if (needed) *needed = out.size() + 1;
        if (buf && cap) std::snprintf(buf, cap, "%s", out.c_str());
        return QSB_OK;
    });
}

int qsb_mc_format_figure_of_merit(qsb_mc* h, double tracking_seconds, char* buf, uint64_t cap)
{
    if (!buf || !cap) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) {
        const double segs = (double)mc.tallies.balanceCumulative[QSB_BAL_NUM_SEGMENTS];
        std::snprintf(buf, cap, "%-25s %12.3e %-25s\n", "Figure Of Merit", tracking_seconds > 0 ? segs / tracking_seconds : 0.0,
                      "[Num Segments / Cycle Tracking Time]");
        return QSB_OK;
    });
}

// Same columns and widths as the reference's per-cycle line (src/Tallies.hh:60-76, src/Tallies.cc:123-144).
int qsb_mc_format_cycle_row(qsb_mc* h, int cycle, const uint64_t row[QSB_BAL_COUNT], double flux,
                            double t_init, double t_track, double t_final, char* buf, uint64_t cap)
{
    if (!row || !buf || !cap) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo&) {
        std::string s;
        char tmp[512];
        if (cycle == 0)
        {
            std::snprintf(tmp, sizeof tmp, "%-8s %12s %12s %12s %12s %12s %12s %12s %12s %12s %12s %12s %12s%14s %14s %14s %14s\n",
                          "cycle", "start", "source", "rr", "split", "absorb", "scatter", "fission", "produce", "collisn",
                          "escape", "census", "num_seg", "scalar_flux", "cycleInit", "cycleTracking", "cycleFinalize");
            s += tmp;
        }
        std::snprintf(tmp, sizeof tmp,
                      "%8i %12llu %12llu %12llu %12llu %12llu %12llu %12llu %12llu %12llu %12llu %12llu %12llu%14e %14e %14e %14e\n",
                      cycle,
                      (unsigned long long)row[QSB_BAL_START], (unsigned long long)row[QSB_BAL_SOURCE],
                      (unsigned long long)row[QSB_BAL_RR], (unsigned long long)row[QSB_BAL_SPLIT],
                      (unsigned long long)row[QSB_BAL_ABSORB], (unsigned long long)row[QSB_BAL_SCATTER],
                      (unsigned long long)row[QSB_BAL_FISSION], (unsigned long long)row[QSB_BAL_PRODUCE],
                      (unsigned long long)row[QSB_BAL_COLLISION], (unsigned long long)row[QSB_BAL_ESCAPE],
                      (unsigned long long)row[QSB_BAL_CENSUS], (unsigned long long)row[QSB_BAL_NUM_SEGMENTS],
                      flux, t_init, t_track, t_final);
        s += tmp;
        std::snprintf(buf, cap, "%s", s.c_str());
        return QSB_OK;
    });
}

} // extern "C"

// Drop-in for the reference's cycleTracking(MonteCarlo*) (src/main.cc:138-307), written purely against the device
// C ABI: host vault in, census + tallies out, copies overlapped with tracking (qsb_stream_*).  Split in two so that a
// multi-rank driver can run its exchange rounds (qsb_track / qsb_send_slab / qsb_put_arrivals) in between:
//   qsb_mc_tracking_begin   cycle_begin + stream_begin(processing vault -> device, census -> processed vault)
//   qsb_mc_tracking_end     stream_end + balance + flux sum into the host model's tallies
extern "C" int qsb_mc_tracking_begin(qsb_mc* h, qsb_ctx* ctx)
{
    if (!ctx) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) {
        auto fail = [&](int rc) { h->error = std::string("device: ") + qsb_last_error(ctx); return rc; };
        int rc;
        mc.timers.trackingStart = nowMicroseconds();
        if ((rc = qsb_cycle_begin(ctx, 0)) != QSB_OK) return fail(rc);
        // the census is delivered straight into the processed vault (page-locked, see MonteCarlo.hh); its size is not
        // known in advance: start from the input size plus headroom, any excess is fetched at the end
        const uint64_t n_in = mc.processing.size();
        uint64_t cap = n_in + n_in / 4 + 65536;
        if (mc.processed.capacity() > cap) cap = mc.processed.capacity();
        mc.processed.clear();
        // growing a page-locked vault costs ~0.5 s/GB (cudaHostAlloc): when it has to grow, take 1.5x so that a
        // population that is still settling does not pay for it again every cycle
        if (cap > mc.processed.capacity()) mc.processed.reserve(cap + cap / 2);
        cap = mc.processed.capacity();
        mc.processed.resize(cap);
        if ((rc = qsb_stream_begin(ctx, mc.processing.data(), n_in, mc.processed.data(), cap)) != QSB_OK) return fail(rc);
        return (int)QSB_OK;
    });
}

extern "C" int qsb_mc_tracking_end(qsb_mc* h, qsb_ctx* ctx)
{
    if (!ctx) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) {
        auto fail = [&](int rc) { h->error = std::string("device: ") + qsb_last_error(ctx); return rc; };
        int rc;
        uint64_t n = 0;
        const uint64_t cap = mc.processed.size();
        if ((rc = qsb_stream_end(ctx, &n)) != QSB_OK) return fail(rc);
        if (n > cap)
        {
            mc.processed.resize(n);
            if ((rc = qsb_get_census_range(ctx, cap, mc.processed.data() + cap, n - cap)) != QSB_OK) return fail(rc);
        }
        mc.processed.resize(n);
        mc.processing.clear();
        mc.residentCensus = false;
        uint64_t bal[QSB_BAL_COUNT];
        double flux = 0.0;
        if ((rc = qsb_get_balance(ctx, bal)) != QSB_OK) return fail(rc);
        if ((rc = qsb_scalar_flux_sum(ctx, &flux)) != QSB_OK) return fail(rc);
        // Tallies::CycleFinalize adds the cycle's flux to the fluence for the CORAL decks (src/Tallies.cc:90-91); the flux
        // lives on the device and is cleared by the next qsb_cycle_begin, so that part of the finalize step runs here
        if (mc.params.simulationParams.coralBenchmark && (rc = qsb_fluence_accumulate(ctx)) != QSB_OK) return fail(rc);
        static const int tracked[] = { QSB_BAL_ABSORB, QSB_BAL_CENSUS, QSB_BAL_ESCAPE, QSB_BAL_COLLISION, QSB_BAL_FISSION,
                                       QSB_BAL_PRODUCE, QSB_BAL_SCATTER, QSB_BAL_NUM_SEGMENTS };
        for (int i : tracked) mc.tallies.balanceTask[i] += bal[i];
        mc.tallies.scalarFluxSum += flux;
        if (mc.timers.trackingStart >= 0.0) mc.timers.add(FastTimers::CycleTracking, nowMicroseconds() - mc.timers.trackingStart, 1);
        mc.timers.trackingStart = -1.0;
        return (int)QSB_OK;
    });
}

extern "C" int qsb_mc_cycle_tracking(qsb_mc* h, qsb_ctx* ctx, qsb_track_stats* stats)
{
    static const bool trace = std::getenv("QSB_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    int rc = qsb_mc_tracking_begin(h, ctx);
    if (rc != QSB_OK) return rc;
    const double t1 = now();
    qsb_track_stats local;
    if ((rc = qsb_track(ctx, stats ? stats : &local)) != QSB_OK)
    {
        if (h) h->error = std::string("device: ") + qsb_last_error(ctx);
        return rc;
    }
    const double t2 = now();
    if (h && h->mc) h->mc->timers.add(FastTimers::CycleTrackingKernel, 1e3 * (double)(stats ? stats->device_ms : local.device_ms),
                                      (stats ? stats->n_launches : local.n_launches));
    rc = qsb_mc_tracking_end(h, ctx);
    if (trace)
        std::fprintf(stderr, "[qsb] cycle_tracking: begin %.2f ms, track %.2f ms (kernel %.2f ms), end %.2f ms\n", t1 - t0, t2 - t1,
                     (double)(stats ? stats->device_ms : local.device_ms), now() - t2);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------------
// The whole cycle with the population resident on the device (include/qsb.h, "device-resident cycles").
// cycleInit (src/main.cc:96-121): the host model keeps what needs the deck or other ranks -- the weight of a source
// particle (src/MC_SourceNow.cc:41-61), the per-cell source counts that follow from it (:72-76), the split / roulette
// factor (src/PopulationControl.cc:32-57) -- and the balance bookkeeping; the per-particle work runs on the device.
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int qsb_mc_cycle_init_resident(qsb_mc* h, qsb_ctx* ctx, qsb_cycle_init_result* result)
{
    if (!ctx) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) {
        auto fail = [&](int rc) { h->error = std::string("device: ") + qsb_last_error(ctx); return rc; };
        int rc;
        mc.timers.clearLastCycle();
        struct InitSection            // cycleInit ends, and the tracking section begins, when this call returns
        {
            FastTimers& t; double t0;
            ~InitSection() { const double t1 = nowMicroseconds(); t.add(FastTimers::CycleInit, t1 - t0, 1); t.trackingStart = t1; }
        } section{ mc.timers, nowMicroseconds() };
        if (!mc.residentCensus)
        {
            // coming from host-side cycles (or the very first cycle): last cycle's census is the processed vault
            if ((rc = qsb_put_census(ctx, mc.processed.data(), mc.processed.size())) != QSB_OK) return fail(rc);
            mc.residentCensusCount = mc.processed.size();
            mc.processed.clear();
            mc.residentCensus = true;
        }
        mc.processing.clear();
        mc.tallies.balanceTask[QSB_BAL_START] = mc.residentCensusCount;          // src/main.cc:106-110
        mc.tallies.scalarFluxSum = 0.0;

        const double weight = buildSourcePlan(mc);
        mc.sourceParticleWeight = weight;
        const size_t nCells = (size_t)mc.image.n_cells;
        const uint64_t nSource = (uint64_t)mc.sourceOffsets[nCells];

        qsb_cycle_init_args a;
        std::memset(&a, 0, sizeof a);
        a.plan_id = mc.sourcePlanId;
        a.source_offsets = mc.sourceOffsets.data();
        // The device keeps and advances its own copy of the running source counts and re-reads ours when the plan id moves
        // (or when the context has no plan yet: a context created at a destroyed one's address must not look familiar, so
        // the arrays are handed over every time -- 8 bytes per cell -- and only the id says whether they are news).
        mc.sourceTallyFlat.resize(nCells);
        {
            size_t flat = 0;
            for (const Domain& d : mc.domain)
                for (int c = 0; c < d.nCells; ++c, ++flat) mc.sourceTallyFlat[flat] = d.sourceTally[c];
        }
        a.source_tally = mc.sourceTallyFlat.data();
        if (mc.devicePlanId != mc.sourcePlanId || mc.devicePlanCtx != (const void*)ctx)
            a.plan_id = ++mc.sourcePlanId;                 // a fresh id: this context has never seen it
        a.source_weight = weight;
        a.e_min = mc.params.simulationParams.eMin; a.e_max = mc.params.simulationParams.eMax;
        a.split_factor = populationControlFactor(mc, mc.residentCensusCount + nSource);
        a.low_weight_cutoff = mc.params.simulationParams.lowWeightCutoff;

        qsb_cycle_init_result local;
        qsb_cycle_init_result* r = result ? result : &local;
        if ((rc = qsb_cycle_init_resident(ctx, &a, r)) != QSB_OK) return fail(rc);
        mc.devicePlanId = mc.sourcePlanId; mc.devicePlanCtx = (const void*)ctx;
        if (r->n_start != mc.residentCensusCount || r->n_source != nSource)
        { h->error = "device census / source count differs from the host model's bookkeeping"; return (int)QSB_ERR_INTERNAL; }
        // keep the host's own running counts in step, so that host-side cycles can take over at any time
        {
            size_t flat = 0;
            for (Domain& d : mc.domain)
                for (int c = 0; c < d.nCells; ++c, ++flat)
                    d.sourceTally[c] += (uint64_t)(mc.sourceOffsets[flat + 1] - mc.sourceOffsets[flat]);
        }
        Balance& bal = mc.tallies.balanceTask;
        bal[QSB_BAL_SOURCE] += r->n_source;
        bal[QSB_BAL_RR] += r->n_rr;
        bal[QSB_BAL_SPLIT] += r->n_split;
        return (int)QSB_OK;
    });
}

extern "C" int qsb_mc_tracking_end_resident(qsb_mc* h, qsb_ctx* ctx)
{
    if (!ctx) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) {
        auto fail = [&](int rc) { h->error = std::string("device: ") + qsb_last_error(ctx); return rc; };
        int rc;
        uint64_t bal[QSB_BAL_COUNT];
        double flux = 0.0;
        uint64_t nCensus = 0;
        if ((rc = qsb_get_balance(ctx, bal)) != QSB_OK) return fail(rc);
        if ((rc = qsb_census_count(ctx, &nCensus)) != QSB_OK) return fail(rc);
        if ((rc = qsb_scalar_flux_sum(ctx, &flux)) != QSB_OK) return fail(rc);
        if (mc.params.simulationParams.coralBenchmark && (rc = qsb_fluence_accumulate(ctx)) != QSB_OK) return fail(rc);
        if (!mc.params.simulationParams.energySpectrum.empty())
        {
            // EnergySpectrum::UpdateSpectrum (src/EnergySpectrum.cc:12-35) over a census that stays on the device
            std::vector<uint64_t> hist(mc.tallies.censusEnergySpectrum.size(), 0);
            if ((rc = qsb_census_energy_spectrum(ctx, hist.data(), hist.size())) != QSB_OK) return fail(rc);
            for (size_t i = 0; i < hist.size(); ++i) mc.tallies.censusEnergySpectrum[i] += hist[i];
            mc.tallies.spectrumDoneThisCycle = true;
        }
        static const int tracked[] = { QSB_BAL_ABSORB, QSB_BAL_CENSUS, QSB_BAL_ESCAPE, QSB_BAL_COLLISION, QSB_BAL_FISSION,
                                       QSB_BAL_PRODUCE, QSB_BAL_SCATTER, QSB_BAL_NUM_SEGMENTS };
        for (int i : tracked) mc.tallies.balanceTask[i] += bal[i];
        mc.tallies.scalarFluxSum += flux;
        mc.residentCensus = true;
        mc.residentCensusCount = nCensus;
        if (mc.timers.trackingStart >= 0.0) mc.timers.add(FastTimers::CycleTracking, nowMicroseconds() - mc.timers.trackingStart, 1);
        mc.timers.trackingStart = -1.0;
        return (int)QSB_OK;
    });
}

extern "C" int qsb_mc_cycle_tracking_resident(qsb_mc* h, qsb_ctx* ctx, qsb_track_stats* stats)
{
    if (!h || !ctx) return QSB_ERR_ARG;
    qsb_track_stats local;
    int rc = qsb_track(ctx, stats ? stats : &local);
    if (rc != QSB_OK)
    {
        h->error = std::string("device: ") + qsb_last_error(ctx);
        return rc;
    }
    if (h->mc) h->mc->timers.add(FastTimers::CycleTrackingKernel, 1e3 * (double)(stats ? stats->device_ms : local.device_ms),
                                 (stats ? stats->n_launches : local.n_launches));
    return qsb_mc_tracking_end_resident(h, ctx);
}

extern "C" int qsb_mc_census_to_host(qsb_mc* h, qsb_ctx* ctx)
{
    if (!ctx) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) {
        if (!mc.residentCensus) return (int)QSB_OK;           // it already is
        uint64_t n = 0;
        int rc = qsb_census_count(ctx, &n);
        if (rc == QSB_OK)
        {
            mc.processed.resize(n);
            rc = qsb_get_census(ctx, mc.processed.data(), n, &n);
        }
        if (rc != QSB_OK) { h->error = std::string("device: ") + qsb_last_error(ctx); return rc; }
        mc.processed.resize(n);
        mc.processing.clear();
        mc.residentCensus = false;
        mc.residentCensusCount = 0;
        return (int)QSB_OK;
    });
}
