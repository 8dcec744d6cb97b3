// capi_host.cc -- extern "C" entry points of the host model (qsb_mc_*), see include/qsb.h.
// Every entry point catches exceptions and reports through the return code + qsb_mc_last_error.
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <new>
#include <string>

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
        else if (k == "nBatches") *out = (int64_t)s.nBatches; else if (k == "batchSize") *out = (int64_t)s.batchSize;
        else if (k == "bTally") *out = s.balanceTallyReplications;
        else if (k == "fTally") *out = s.fluxTallyReplications;
        else if (k == "cycle") *out = mc.cycle;
        else if (k == "nDomains") *out = (int64_t)mc.domain.size();
        else if (k == "nCells") *out = mc.image.n_cells;
        else if (k == "nProcessing") *out = (int64_t)mc.processing.size();
        else if (k == "nProcessed") *out = (int64_t)mc.processed.size();
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
    return guarded(h, [&](MonteCarlo& mc) { cycleInit(mc); return QSB_OK; });
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
        Balance r; double f = 0;
        cycleFinalize(mc, r, f);
        if (row) std::memcpy(row, r.v, sizeof(r.v));
        if (flux) *flux = f;
        return QSB_OK;
    });
}

int qsb_mc_cumulative_balance(qsb_mc* h, uint64_t out[QSB_BAL_COUNT])
{
    if (!out) return QSB_ERR_ARG;
    return guarded(h, [&](MonteCarlo& mc) { std::memcpy(out, mc.tallies.balanceCumulative.v, sizeof(uint64_t) * QSB_BAL_COUNT); return QSB_OK; });
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
        uint64_t bal[QSB_BAL_COUNT];
        double flux = 0.0;
        if ((rc = qsb_get_balance(ctx, bal)) != QSB_OK) return fail(rc);
        if ((rc = qsb_scalar_flux_sum(ctx, &flux)) != QSB_OK) return fail(rc);
        static const int tracked[] = { QSB_BAL_ABSORB, QSB_BAL_CENSUS, QSB_BAL_ESCAPE, QSB_BAL_COLLISION, QSB_BAL_FISSION,
                                       QSB_BAL_PRODUCE, QSB_BAL_SCATTER, QSB_BAL_NUM_SEGMENTS };
        for (int i : tracked) mc.tallies.balanceTask[i] += bal[i];
        mc.tallies.scalarFluxSum += flux;
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
    rc = qsb_mc_tracking_end(h, ctx);
    if (trace)
        std::fprintf(stderr, "[qsb] cycle_tracking: begin %.2f ms, track %.2f ms (kernel %.2f ms), end %.2f ms\n", t1 - t0, t2 - t1,
                     (double)(stats ? stats->device_ms : local.device_ms), now() - t2);
    return rc;
}
