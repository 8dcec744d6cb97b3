"""Cycle driver: the reference's main loop (src/main.cc:38-121,138-324) for one process per GPU.

Single GPU: cycle_init (host) -> qsb_mc_cycle_tracking (device) -> cycle_finalize (host).
Several GPUs: the spatial domain decomposition gives each rank one domain; boundary-crossing particles
leave the tracking kernel in per-peer slabs and are exchanged between ranks with torch.distributed
(NCCL send/recv over NVLink for device slabs, gloo for the CPU protocol tests); termination is the
reference's "gains == losses" test reduced to "nobody sent anything this round"
(src/MC_Particle_Buffer.cc:601-618).  torch is plumbing only: device memory views and collectives.
"""
import time

import numpy as np

from . import _capi, device as device_mod, host as host_mod
from ._capi import BAL, BAL_COUNT, EXCHANGE_DTYPE, PARTICLE_DTYPE

RECORD_BYTES = EXCHANGE_DTYPE.itemsize


class _DeviceMemory:
    """__cuda_array_interface__ shim so torch can view a slab owned by the qsb_ctx without copying."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class DeviceBackend:
    """tracking backend over a qsb_ctx (the product path)."""

    def __init__(self, ctx, torch_device):
        self.ctx, self.torch_device = ctx, torch_device

    @property
    def peer_mode(self):
        return self.ctx.peer_mode

    def begin(self, vault):
        self.ctx.cycle_begin()
        self.ctx.put_particles(vault)

    def begin_streamed(self, mc):
        """the host model's processing vault is streamed to the device under the first tracking round and the
        census streamed back into its processed vault (qsb_mc_tracking_begin)."""
        mc.tracking_begin(self.ctx)

    def end_streamed(self, mc):
        mc.tracking_end(self.ctx)

    def track(self):
        stats = self.ctx.track()
        self.device_ms = getattr(self, "device_ms", 0.0) + stats.device_ms
        return stats

    def send_counts(self):
        return self.ctx.send_counts().astype(np.int64)

    def send_tensor(self, peer, n):
        import torch
        ptr, have = self.ctx.send_slab(peer)
        assert have == n
        return torch.as_tensor(_DeviceMemory(ptr, n * RECORD_BYTES), device=self.torch_device)

    def recv_tensor(self, n):
        import torch
        return torch.empty(n * RECORD_BYTES, dtype=torch.uint8, device=self.torch_device)

    def put_arrivals(self, tensor, n):
        import torch
        torch.cuda.synchronize(self.torch_device)
        self.ctx.put_arrivals(tensor.data_ptr(), n)

    def clear_sends(self):
        self.ctx.clear_sends()

    def results(self):
        return self.ctx.get_census(), self.ctx.get_balance(), self.ctx.scalar_flux_sum()


def connect_peers(ctx, dist, rank, world, torch_device, watchdog_seconds=0.0):
    """Wire the ranks' exchange rings together over NVLink: every rank exports its ring as a CUDA IPC handle, the handles
    are all-gathered, every rank maps the others' rings (qsb_peer_connect).  All ranks agree on the outcome: True = the
    tracking kernels exchange boundary particles themselves (one launch per cycle), False = NCCL rounds (exchange_rounds)."""
    import torch
    if world < 2 or world > ctx.MAX_PEERS:
        return False
    ok, mine, cap = 1, bytes(ctx.PEER_HANDLE_BYTES), 0
    try:
        mine, cap = ctx.peer_export()
    except Exception:
        ok = 0
    caps = torch.tensor([cap, -cap], dtype=torch.int64, device=torch_device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(caps, op=dist.ReduceOp.MAX)
    if int(caps[0].item()) != -int(caps[1].item()):      # max != min: the ranks' vaults differ in capacity
        ok = 0
    is_cuda = dist.get_backend() == "nccl"
    dev = torch_device if is_cuda else "cpu"
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        return False
    local = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    handles = b"".join(bytes(t.cpu().numpy().tobytes()) for t in gathered)
    try:
        ctx.peer_connect(handles, watchdog_seconds)
        import os
        if os.environ.get("QSB_DEBUG_PEER_MAP_ONLY"):
            ctx.peer_mode = False
            return False
    except Exception:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        ctx._lib.qsb_peer_disconnect(ctx._h)
        ctx.peer_mode = False
        return False
    return True


def exchange_rounds(backend, dist, rank, world, max_rounds=100000):
    """Track to exhaustion, swap boundary particles, repeat until no rank sent anything.
    Returns (rounds, records sent by this rank).  With the rings wired over NVLink (connect_peers) the kernels do the
    exchange and the termination test themselves: one launch, one "round"."""
    import torch
    sent_total, rounds = 0, 0
    if getattr(backend, "peer_mode", False):
        stats = backend.track()
        return 1, int(stats.n_sent)
    while True:
        backend.track()
        rounds += 1
        if world == 1:
            return rounds, 0
        counts = backend.send_counts()
        counts[rank] = 0
        is_cuda = dist.get_backend() == "nccl"
        dev = backend.torch_device if is_cuda else "cpu"
        send_counts = torch.as_tensor(counts, dtype=torch.int64, device=dev)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts)
        total = send_counts.sum().reshape(1).clone()
        dist.all_reduce(total)
        recv = recv_counts.cpu().numpy()
        if int(total.item()) == 0:
            return rounds, sent_total
        n_in = int(recv.sum())
        inbox = backend.recv_tensor(n_in)
        ops, offset, keep = [], 0, []
        for peer in range(world):
            if peer == rank:
                continue
            if recv[peer]:
                nbytes = int(recv[peer]) * RECORD_BYTES
                ops.append(dist.P2POp(dist.irecv, inbox[offset:offset + nbytes], peer))
                offset += nbytes
            if counts[peer]:
                t = backend.send_tensor(peer, int(counts[peer]))
                keep.append(t)
                ops.append(dist.P2POp(dist.isend, t, peer))
        if ops:
            for work in dist.batch_isend_irecv(ops):
                work.wait()
        sent_total += int(counts.sum())
        backend.clear_sends()
        if n_in:
            backend.put_arrivals(inbox, n_in)
        if rounds >= max_rounds:
            raise RuntimeError("particle exchange did not terminate")


def bind_to_gpu_numa_node(device):
    """Run this process (and therefore first-touch its page-locked vaults) on the CPUs of the NUMA node the GPU hangs off, so
    that the streamed host vaults do not cross the socket interconnect on their way to and from the device.  Several ranks
    of one node otherwise land wherever the scheduler puts them.  Best effort: returns (node, why): node is None when nothing
    was bound, and `why` says what stood in the way (bench.py prints it as numa_node_rank0 / numa_note)."""
    import os
    import subprocess
    if os.environ.get("QSB_NO_NUMA_BIND"):
        return None, "QSB_NO_NUMA_BIND set"
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device)],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=20).stdout.strip().splitlines()[0].strip()
        dom, rest = out.split(":", 1)                      # "00000000:1B:00.0" -> "0000:1b:00.0"
        bdf = ("%04x:%s" % (int(dom, 16), rest)).lower()
    except Exception as e:
        return None, "nvidia-smi gave no PCI bus id (%s)" % type(e).__name__
    try:
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
    except Exception:
        return None, "no /sys/bus/pci/devices/%s/numa_node in this container" % bdf
    if node < 0:
        try:
            n_nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
        except Exception:
            n_nodes = 0
        return None, "the platform reports numa_node = -1 for %s (%d NUMA node(s) visible): nothing to bind to" % (bdf, n_nodes)
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None, "none of NUMA node %d's CPUs is in this process's affinity mask (cgroup cpuset)" % node
        os.sched_setaffinity(0, allowed)
        return node, "bound to %d CPUs of node %d" % (len(allowed), node)
    except Exception as e:
        return None, "could not read / apply the CPU list of node %d (%s)" % (node, type(e).__name__)


class Simulation:
    """One rank of a run: host model + device context + the cycle loop."""

    def __init__(self, argv, rank=0, world=1, device=0, validation=True, dist=None, particle_capacity=0, send_capacity=0,
                 make_backend=None, resident=False):
        """make_backend(mc) -> tracking backend; None = the device (the product path).  Tests of the exchange
        protocol pass a CPU stand-in so that the N-rank logic runs under gloo on a machine without GPUs.
        resident: keep the particle population on the device from cycle to cycle -- cycleInit's per-particle work (source,
        population control, roulette) runs there too (qsb_mc_cycle_init_resident); only tallies cross PCIe."""
        self.rank, self.world, self.dist = rank, world, dist
        self.resident = bool(resident) and make_backend is None
        self.torch_device = "cuda:%d" % device
        self.numa_node, self.numa_note = bind_to_gpu_numa_node(device) if (make_backend is None and world > 1) else (None, "single rank: not bound")
        self.mc = host_mod.MonteCarlo(argv, rank, world, allreduce=self._allreduce if world > 1 else None)
        self.ctx = None
        if make_backend is not None:
            self.backend = make_backend(self.mc)
            return
        n_particles = self.mc.get_int("nParticles")
        if particle_capacity == 0:
            per_rank = (n_particles + world - 1) // world
            particle_capacity = int(per_rank * (3 + 2 * self.mc.get_double("max_nu_bar"))) + (1 << 16)
        self.ctx = device_mod.DeviceContext(self.mc.image, self.mc.get_double("dt"), device=device, validation=validation,
                                            particle_capacity=particle_capacity, send_capacity=send_capacity)
        self.backend = DeviceBackend(self.ctx, self.torch_device)
        # boundary particles: device-to-device rings over NVLink where available ("peer"), else NCCL send/recv rounds
        import os
        want = os.environ.get("QSB_EXCHANGE", "peer")
        self.exchange = "nccl"
        if world > 1 and want == "peer" and dist is not None and dist.get_backend() == "nccl":
            if connect_peers(self.ctx, dist, rank, world, self.torch_device, float(os.environ.get("QSB_PEER_WATCHDOG_S", "0"))):
                self.exchange = "peer"

    def _allreduce(self, arr, op="sum"):
        import torch
        is_cuda = self.dist.get_backend() == "nccl"
        view = arr.view(np.int64) if arr.dtype == np.uint64 else arr
        t = torch.from_numpy(view.copy())
        if is_cuda:
            t = t.to(self.torch_device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        view[:] = t.cpu().numpy()

    def report(self, tracking_seconds=None):
        """the reference's closing report (src/main.cc:66-94): timer table + figure of merit (MC_Fast_Timer), spectrum file,
        CORAL self checks; text on rank 0, every rank calls it (the reports reduce over ranks)."""
        timers = self.mc.timer_report()
        self.mc.write_energy_spectrum()
        fluence = self.ctx.get_fluence() if self.ctx is not None else None
        text, passed = self.mc.coral_benchmark_report(fluence)
        return timers + text, passed

    def _feed_exchange_timers(self, t_track, rounds):
        """several ranks: the tracking section is driven from here (exchange rounds), so its kernel / exchange split is ours
        to report: cycleTracking_Kernel = CUDA-event time of the launches, cycleTracking_MPI = the rest of the section."""
        kernel_us = 1e3 * getattr(self.backend, "device_ms", 0.0)
        self.mc.timer_add("cycleTracking_Kernel", kernel_us, rounds)
        self.mc.timer_add("cycleTracking_MPI", max(0.0, 1e6 * t_track - kernel_us), rounds)
        self.mc.timer_add("cycleTracking_Test_Done", 0.0, rounds)

    def cycle(self):
        """one cycle; returns (global balance row, global flux sum, timings dict)."""
        t0 = time.perf_counter()
        info = {}
        if self.resident:
            res = self.mc.cycle_init_resident(self.ctx)
            t1 = time.perf_counter()
            if self.world == 1:
                stats = self.mc.cycle_tracking_resident(self.ctx)
                info.update(device_ms=stats.device_ms, launches=stats.n_launches, rounds=1)
            else:
                self.backend.device_ms = 0.0
                rounds, sent = exchange_rounds(self.backend, self.dist, self.rank, self.world)
                self.mc.tracking_end_resident(self.ctx)
                self._feed_exchange_timers(time.perf_counter() - t1, rounds)
                info.update(rounds=rounds, sent=sent)
            t2 = time.perf_counter()
            row, flux = self.mc.cycle_finalize()
            t3 = time.perf_counter()
            info.update(t_init=t1 - t0, t_track=t2 - t1, t_final=t3 - t2, init_device_ms=res.device_ms, n_processing=int(res.n_processing))
            return row, flux, info
        self.mc.cycle_init()
        t1 = time.perf_counter()
        if self.world == 1 and self.ctx is not None:
            stats = self.mc.cycle_tracking(self.ctx)
            info.update(device_ms=stats.device_ms, launches=stats.n_launches, rounds=1)
        elif hasattr(self.backend, "begin_streamed"):
            self.backend.device_ms = 0.0
            self.backend.begin_streamed(self.mc)
            rounds, sent = exchange_rounds(self.backend, self.dist, self.rank, self.world)
            self.backend.end_streamed(self.mc)
            self._feed_exchange_timers(time.perf_counter() - t1, rounds)
            info.update(rounds=rounds, sent=sent)
        else:
            vault = self.mc.processing()
            self.backend.begin(vault)
            rounds, sent = exchange_rounds(self.backend, self.dist, self.rank, self.world)
            census, balance, flux_sum = self.backend.results()
            self.mc.set_tracking_result(census, balance, flux_sum)
            self.mc.timer_add("cycleTracking", 1e6 * (time.perf_counter() - t1), 1)
            info.update(rounds=rounds, sent=sent)
        t2 = time.perf_counter()
        row, flux = self.mc.cycle_finalize()
        t3 = time.perf_counter()
        info.update(t_init=t1 - t0, t_track=t2 - t1, t_final=t3 - t2)
        return row, flux, info

    def close(self):
        if self.ctx is not None:
            self.ctx.close()
        self.mc.close()


def run_resident_workload(argv, rank, world, local_rank, dist, validation, warmup, steps, env=None):
    """A few whole cycles of one deck with the population resident on the device (cycleInit on the GPU, one tracking launch
    per cycle, peer exchange between GPUs): the figure of merit of a secondary workload in bench.py's `workloads` block and
    of the validation build.  Timing as for the headline `value`: one GPU -> CUDA events around the tracking kernel on its
    own stream; several GPUs -> host clock between device synchronisations + barriers around the exchange, max over ranks.
    Returns a dict (identical on every rank)."""
    import os
    import torch
    saved = {}
    for k, v in (env or {}).items():
        saved[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        sim = Simulation(argv, rank, world, device=local_rank, validation=validation, dist=dist if world > 1 else None, resident=True)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    dev = "cuda:%d" % local_rank

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    track_s = init_s = dev_track_s = 0.0
    segments = 0
    rows = []
    launches0 = 0
    for step in range(warmup + steps):
        timed = step >= warmup
        if step == warmup:
            launches0 = sim.ctx.launch_count()
        res = sim.mc.cycle_init_resident(sim.ctx)
        barrier()
        t0 = time.perf_counter()
        sim.backend.device_ms = 0.0
        if world == 1:
            stats = sim.mc.cycle_tracking_resident(sim.ctx)
            step_dev = stats.device_ms * 1e-3
        else:
            exchange_rounds(sim.backend, dist, rank, world)
            torch.cuda.synchronize()
            step_dev = sim.backend.device_ms * 1e-3
        t1 = time.perf_counter()
        if world > 1:
            sim.mc.tracking_end_resident(sim.ctx)
        row, flux = sim.mc.cycle_finalize()
        rows.append([int(v) for v in row])
        if timed:
            track_s += step_dev if world == 1 else (t1 - t0)
            dev_track_s += step_dev
            init_s += res.device_ms * 1e-3
            segments += int(row[BAL["num_segments"]])
    launches = sim.ctx.launch_count() - launches0
    t = torch.tensor([track_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    track_max = float(t.cpu()[0])
    cum = sim.mc.cumulative_balance()
    gains = int(cum[BAL["start"]] + cum[BAL["source"]] + cum[BAL["produce"]] + cum[BAL["split"]])
    losses = int(cum[BAL["absorb"]] + cum[BAL["census"]] + cum[BAL["escape"]] + cum[BAL["rr"]] + cum[BAL["fission"]])
    out = {"value": segments / track_max if track_max > 0 else 0.0, "unit": "segments/s", "steps": steps, "warmup": warmup,
           "ms_per_step": 1e3 * track_max / max(steps, 1), "segments_per_step": segments // max(steps, 1),
           "track_kernel_ms_rank0": 1e3 * dev_track_s / max(steps, 1), "cycle_init_kernel_ms_rank0": 1e3 * init_s / max(steps, 1),
           "kernels": "validation" if validation else "fast", "gpu_launches": launches, "conserved": gains == losses,
           "exchange": getattr(sim, "exchange", "none") if world > 1 else "none", "last_row": rows[-1],
           "cells_rank0": int(sim.mc.image.n_cells), "groups": int(sim.mc.image.n_groups)}
    sim.close()
    return out


def run_benchmark(args, warmup, rank, world, local_rank, workloads, grid_ladder, deck_argv, ClockSampler, dist=None):
    """bench.py's B200 arm.  Times, per step, (a) the device-resident tracking (CUDA events inside qsb_track)
    and (b) the host-buffer drop-in call, and returns totals reduced over ranks (max of times, sum of segments).
    `dist`: the initialised torch.distributed module when world > 1 (bench.py owns the process group)."""
    import tempfile
    import torch

    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    w = dict(workloads[args.workload])
    if args.scale != 1.0:
        k = args.scale ** (1.0 / 3.0)
        w["n"] = max(4, int(round(w["n"] * k)))
        w["particles"] = int(w["particles"] * (w["n"] / workloads[args.workload]["n"]) ** 3)
    grid = grid_ladder[world]
    tmp = tempfile.mkdtemp(prefix="qsb_bench_")
    argv = deck_argv(w, grid, tmp, warmup + args.steps)

    sim = Simulation(argv, rank, world, device=local_rank, validation=not args.fast, dist=dist if world > 1 else None)
    mc, ctx = sim.mc, sim.ctx

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    kernel_s = e2e_s = device_s = host_s = host_init_s = host_final_s = 0.0
    sent_total = 0
    segments = 0
    h2d = d2h = 0
    launches0 = 0
    rows = []
    sampler = ClockSampler(local_rank)
    peer_diag = {}
    for step in range(warmup + args.steps):
        timed = step >= warmup
        if step == warmup:
            barrier()
            sampler.start()
            launches0 = ctx.launch_count()
        t_ci = time.perf_counter()
        mc.cycle_init()
        if timed:
            host_init_s += time.perf_counter() - t_ci
        n_in = mc.get_int("nProcessing")
        # (a) `value`: the vault already resident in HBM when the timed region starts.  A first pass of the cycle whose
        #     census is discarded (pass (b) redoes the cycle from the same host vault and is the one whose results are kept).
        #     One GPU: CUDA events around the tracking kernel (qsb_track's own stream).  Several GPUs: the exchange rounds
        #     (kernels + NCCL send/recv + termination allreduce, what the reference's cycleTracking timer covers) between two
        #     device synchronisations + barriers, max over ranks.
        vault = mc.processing()
        sim.backend.begin(vault)
        del vault
        barrier()
        ta = time.perf_counter()
        sim.backend.device_ms = 0.0
        _, n_sent = exchange_rounds(sim.backend, dist, rank, world)
        torch.cuda.synchronize()
        tb = time.perf_counter()
        if timed:
            sent_total += n_sent
            if world > 1 and getattr(sim, "exchange", "") == "peer":
                d = ctx.peer_diagnostics()
                for k in ("first_idle_ns", "done_ns", "send_cycles", "send_calls", "startup_wait_ns", "deposited", "bulk_done_ns"):
                    peer_diag[k] = peer_diag.get(k, 0) + d[k]
        step_kernel_s = sim.backend.device_ms * 1e-3 if world == 1 else tb - ta
        if timed:
            device_s += sim.backend.device_ms * 1e-3
            host_s += tb - ta
        # (b) `e2e`: host buffers in, host buffers out -- the drop-in call (one GPU) / its two halves around the exchange
        #     rounds (several GPUs), wall clock, copies of the vaults included (streamed under the tracking)
        barrier()
        t0 = time.perf_counter()
        if getattr(args, "resident_only", 0):
            census, balance, flux_sum = sim.backend.results()       # profiling runs: keep pass (a), no streamed pass
            mc.set_tracking_result(census, balance, flux_sum)
        elif world == 1:
            stats = mc.cycle_tracking(ctx)
        else:
            sim.backend.begin_streamed(mc)
            exchange_rounds(sim.backend, dist, rank, world)
            sim.backend.end_streamed(mc)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        step_e2e_s = t1 - t0
        n_census = mc.get_int("nProcessed")
        t_cf = time.perf_counter()
        row, flux = mc.cycle_finalize()     # global row (allreduced)
        if timed:
            host_final_s += time.perf_counter() - t_cf
        rows.append([int(v) for v in row])
        if timed:
            kernel_s += step_kernel_s
            e2e_s += step_e2e_s
            segments += int(row[BAL["num_segments"]])
            h2d += n_in * PARTICLE_DTYPE.itemsize
            d2h += n_census * PARTICLE_DTYPE.itemsize + BAL_COUNT * 8 + 8
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0

    # ---- the WHOLE cycle (cycleInit + cycleTracking + cycleFinalize, what a run's wall clock is made of) in the two
    #      arrangements: host-staged (above: host cycleInit, vaults through PCIe inside the drop-in call) and resident
    #      (the same simulation continued with the population left on the device and cycleInit's per-particle work done
    #      there, qsb_mc_cycle_init_resident).  Reported next to the FOM, never instead of it.
    whole = {"host_staged": {"cycle_init_ms": 1e3 * host_init_s / max(args.steps, 1), "cycle_tracking_ms": 1e3 * e2e_s / max(args.steps, 1),
                             "cycle_finalize_ms": 1e3 * host_final_s / max(args.steps, 1)}}
    import os
    want_resident = os.environ.get("QSB_BENCH_RESIDENT", "1")          # "0": skip; "force": also in --resident-only (ncu) runs
    if want_resident != "0" and (not getattr(args, "resident_only", 0) or want_resident == "force"):
        try:
            n_res = 20                                       # ~0.4 s of back-to-back cycles at benchmark size
            r_init = r_track = r_final = r_init_dev = r_track_dev = 0.0
            r_segments = 0
            r_carried = 0
            r_launches0 = 0
            r_sampler = ClockSampler(local_rank, 50)           # back-to-back cycles keep the GPU busy: clocks under sustained load
            for k in range(1 + n_res):                       # the first one moves the census to the device: not timed
                barrier()
                if k == 1:
                    r_launches0 = ctx.launch_count()
                    r_sampler.start()
                t0 = time.perf_counter()
                res = mc.cycle_init_resident(ctx)
                t1 = time.perf_counter()
                if world == 1:
                    stats = mc.cycle_tracking_resident(ctx)
                    track_dev_ms = stats.device_ms
                else:
                    sim.backend.device_ms = 0.0
                    exchange_rounds(sim.backend, dist, rank, world)
                    mc.tracking_end_resident(ctx)
                    track_dev_ms = sim.backend.device_ms
                t2 = time.perf_counter()
                row, flux = mc.cycle_finalize()
                t3 = time.perf_counter()
                rows.append([int(v) for v in row])
                if k >= 1:
                    r_carried += int(res.n_start)
                    r_init += t1 - t0; r_track += t2 - t1; r_final += t3 - t2
                    r_init_dev += res.device_ms * 1e-3; r_track_dev += track_dev_ms * 1e-3
                    r_segments += int(row[BAL["num_segments"]])
            r_clocks = r_sampler.stop()
            rt = torch.tensor([r_init, r_track, r_final, r_init + r_track + r_final], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(rt, op=dist.ReduceOp.MAX)
            r_init, r_track, r_final, r_total = (float(v) for v in rt.cpu())
            whole["resident"] = {"cycles": n_res, "cycle_init_ms": 1e3 * r_init / n_res, "cycle_tracking_ms": 1e3 * r_track / n_res,
                                 "cycle_finalize_ms": 1e3 * r_final / n_res, "ms_per_cycle": 1e3 * r_total / n_res,
                                 "segments_per_s_whole_cycle": r_segments / r_total if r_total > 0 else 0.0,
                                 "segments_per_s_tracking": r_segments / r_track if r_track > 0 else 0.0,
                                 "cycle_init_kernel_ms_rank0": 1e3 * r_init_dev / n_res, "track_kernel_ms_rank0": 1e3 * r_track_dev / n_res,
                                 "carried_per_cycle": r_carried / n_res,
                                 "gpu_launches": ctx.launch_count() - r_launches0,
                                 "pcie_bytes_per_cycle": BAL_COUNT * 8 + 8 + 48, "clocks": r_clocks,
                                 "note": "wall clock per rank around each stage, max over ranks; population resident in HBM, source + population "
                                         "control + low-weight roulette in one kernel on the device"}
            if os.environ.get("QSB_BENCH_RESIDENT_CHECK") and world == 1:
                # diagnostic: the SAME population taken back through the host path (census -> host, host cycleInit, vault
                # uploaded, tracking kernel timed alone) -- separates "the resident vault is laid out differently" from
                # "the population of these later cycles tracks differently"
                check = []
                for _ in range(3):
                    mc.census_to_host(ctx)
                    mc.cycle_init()
                    sim.backend.begin(mc.processing())
                    st = ctx.track()
                    check.append(st.device_ms)
                    census, balance, flux_sum = sim.backend.results()
                    mc.set_tracking_result(census, balance, flux_sum)
                    row, _ = mc.cycle_finalize()
                    rows.append([int(v) for v in row])
                    # and one resident cycle on top of that host-built census
                    res = mc.cycle_init_resident(ctx)
                    st = mc.cycle_tracking_resident(ctx)
                    check.append(-st.device_ms)
                    row, _ = mc.cycle_finalize()
                    rows.append([int(v) for v in row])
                whole["resident"]["check_track_kernel_ms_host_built_vault_then_resident(negative)"] = check
        except Exception as e:          # an extra; the FOM line above never depends on it
            whole["resident"] = {"error": "%s: %s" % (type(e).__name__, e)}
    hs = whole["host_staged"]
    hs["ms_per_cycle"] = hs["cycle_init_ms"] + hs["cycle_tracking_ms"] + hs["cycle_finalize_ms"]
    hs["segments_per_s_whole_cycle"] = (segments / (hs["ms_per_cycle"] * 1e-3 * max(args.steps, 1))) if hs["ms_per_cycle"] > 0 else 0.0

    # reduce over ranks: max of times, segments already global
    times = torch.tensor([kernel_s, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    kernel_max, e2e_max = (float(v) for v in times.cpu())
    # conservation identity of the reference's MissingParticleTest (src/CoralBenchmark.cc:152-174)
    cum = mc.cumulative_balance()
    gains = int(cum[BAL["start"]] + cum[BAL["source"]] + cum[BAL["produce"]] + cum[BAL["split"]])
    losses = int(cum[BAL["absorb"]] + cum[BAL["census"]] + cum[BAL["escape"]] + cum[BAL["rr"]] + cum[BAL["fission"]])
    n = w["n"]
    config = {"workload": "%s weak-scaled: %dx%dx%d cells and %d particles per GPU, %s domain grid" % (
                  args.workload, n, n, n, w["particles"], "x".join(str(g) for g in grid)),
              "deck": w["deck"], "cells_per_gpu": n ** 3, "particles_per_gpu": w["particles"], "domain_grid": list(grid),
              "kernels": "fast" if args.fast else "validation", "timing": "inputs larger than L2 (vault %.0f MB, mesh %.0f MB per GPU); value: %s; e2e: wall clock around the "
              "drop-in call with host vaults" % (w["particles"] * 160 / 1e6, n ** 3 * 1.2e-3,
              "CUDA events on the tracking stream" if world == 1 else "host clock between device syncs + barriers around the exchange rounds, max over ranks"),
              "scale": args.scale, "exchange": getattr(sim, "exchange", "none") if world > 1 else "none", "numa_node_rank0": sim.numa_node, "numa_note": sim.numa_note}
    out = {"segments_total": segments, "kernel_seconds_max": kernel_max, "e2e_seconds_max": e2e_max,
           "segments_rank0": segments / world, "kernel_seconds_rank0": kernel_s, "config": config, "clocks": clocks,
           "h2d_bytes_per_step": h2d // max(args.steps, 1), "d2h_bytes_per_step": d2h // max(args.steps, 1),
           "gpu_launches": launches, "traffic": None, "whole_cycle": whole,
           "tracking_ms_per_step_rank0": {"boundary_particles_sent": sent_total // max(args.steps, 1),
                                          "cuda_events_on_kernel_stream": 1e3 * device_s / max(args.steps, 1),
                                          "host_clock_around_call": 1e3 * host_s / max(args.steps, 1)},
           "balance_check": {"gains": gains, "losses": losses, "conserved": gains == losses, "last_row": rows[-1]}}
    if world > 1:
        # every rank's own numbers, so that the slowest rank and the termination tail are visible in the line (per timed step)
        k = max(args.steps, 1)
        mine = {"rank": rank, "kernel_ms": 1e3 * device_s / k, "host_clock_ms": 1e3 * host_s / k, "e2e_ms": 1e3 * e2e_s / k,
                "boundary_particles_sent": sent_total // k}
        if peer_diag:
            mine.update(first_idle_ms=peer_diag["first_idle_ns"] * 1e-6 / k, done_ms=peer_diag["done_ns"] * 1e-6 / k,
                        tail_ms=(peer_diag["done_ns"] - peer_diag["first_idle_ns"]) * 1e-6 / k,
                        startup_wait_ms=peer_diag["startup_wait_ns"] * 1e-6 / k, own_queue_empty_ms=peer_diag.get("bulk_done_ns", 0) * 1e-6 / k,
                        send_mcycles_all_warps=peer_diag["send_cycles"] * 1e-6 / k, send_passes=peer_diag["send_calls"] // k)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        out["per_rank"] = gathered
    sim.close()
    return out
