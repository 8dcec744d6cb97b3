// qs_main.cc -- the `qs_b200` executable: the reference's command line, cycle loop and report
// (src/main.cc:38-94) over the C ABI of libqsb.so, single GPU.  Multi-GPU runs are driven by
// quicksilver_b200/driver.py (one process per GPU under torch.distributed).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../../include/qsb.h"

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv)
{
    qsb_mc* mc = nullptr;
    if (qsb_mc_create(argc, argv, 0, 1, &mc) != QSB_OK)
    {
        std::fprintf(stderr, "%s\n", qsb_mc_last_error(nullptr));
        return 2;
    }
    uint64_t need = 0;
    qsb_mc_print_parameters(mc, nullptr, 0, &need);
    std::vector<char> text(need);
    qsb_mc_print_parameters(mc, text.data(), need, nullptr);
    std::printf("%s", text.data());

    qsb_image image;
    qsb_mc_get_image(mc, &image);
    int64_t n_steps = 0, n_particles = 0;
    double dt = 0, nu_bar = 0;
    qsb_mc_get_int(mc, "nSteps", &n_steps);
    qsb_mc_get_int(mc, "nParticles", &n_particles);
    qsb_mc_get_double(mc, "dt", &dt);
    qsb_mc_get_double(mc, "max_nu_bar", &nu_bar);

    qsb_options opt = {};
    const char* fast = std::getenv("QSB_FAST");
    opt.validation = (fast && fast[0] == '1') ? 0 : 1;
    opt.particle_capacity = (uint64_t)(n_particles * 6 + 65536);
    qsb_ctx* ctx = nullptr;
    if (qsb_create(0, &image, dt, &opt, &ctx) != QSB_OK)
    {
        std::fprintf(stderr, "qsb_create: %s\n", qsb_last_error(nullptr));
        return 3;
    }

    // Where the population lives between cycles.  Host (the reference's arrangement: cycleInit on the host, vaults through
    // PCIe every cycle; libm source -> the reference binary's table bit for bit in the validation build) or resident on the
    // device (cycleInit's per-particle work on the GPU too; same streams, log/sin/cos of the source rounded differently in
    // the last bit).  Default: host for the validation build, resident for the fast build; QSB_RESIDENT=0/1 overrides.
    const char* res_env = std::getenv("QSB_RESIDENT");
    const bool resident = res_env ? res_env[0] == '1' : opt.validation == 0;

    int64_t cycle_timers = 0;
    qsb_mc_get_int(mc, "cycleTimers", &cycle_timers);
    double t_track_total = 0;
    for (int cycle = 0; cycle < n_steps; ++cycle)
    {
        const double t0 = now();
        const int rc_init = resident ? qsb_mc_cycle_init_resident(mc, ctx, nullptr) : qsb_mc_cycle_init(mc);
        if (rc_init != QSB_OK) { std::fprintf(stderr, "%s\n", qsb_mc_last_error(mc)); return 4; }
        const double t1 = now();
        qsb_track_stats stats;
        const int rc_track = resident ? qsb_mc_cycle_tracking_resident(mc, ctx, &stats) : qsb_mc_cycle_tracking(mc, ctx, &stats);
        if (rc_track != QSB_OK) { std::fprintf(stderr, "%s\n", qsb_mc_last_error(mc)); return 5; }
        const double t2 = now();
        uint64_t row[QSB_BAL_COUNT]; double flux = 0;
        qsb_mc_cycle_finalize(mc, row, &flux);
        const double t3 = now();
        char line[2048];
        qsb_mc_format_cycle_row(mc, cycle, row, flux, t1 - t0, t2 - t1, t3 - t2, line, sizeof line);
        std::printf("%s", line);
        t_track_total += t2 - t1;
        if (cycle_timers)                                          // Last_Cycle_Report (src/main.cc:61-65)
        {
            char report[4096];
            qsb_mc_format_timer_report(mc, 1, report, sizeof report, nullptr);
            std::printf("%s", report);
        }
    }
    // gameOver (src/main.cc:87-94): timer table + figure of merit (src/MC_Fast_Timer.cc:58-105), spectrum file; then
    // coralBenchmarkCorrectness (src/main.cc:73, src/CoralBenchmark.cc)
    {
        char timers[4096];
        qsb_mc_format_timer_report(mc, 0, timers, sizeof timers, nullptr);
        std::printf("%s", timers);
        qsb_mc_write_energy_spectrum(mc);
        std::vector<double> fluence((size_t)image.n_cells, 0.0);
        qsb_get_fluence(ctx, fluence.data());
        std::vector<char> report(8192);
        qsb_mc_coral_benchmark_report(mc, fluence.data(), fluence.size(), report.data(), report.size(), nullptr, nullptr);
        std::printf("%s", report.data());
    }
    (void)t_track_total;
    qsb_destroy(ctx);
    qsb_mc_destroy(mc);
    return 0;
}
