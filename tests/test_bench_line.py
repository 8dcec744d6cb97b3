"""bench.py's JSON line: the contract keys are assembled from the driver's result without a GPU (the driver's measurement is
replaced by canned numbers), and the reference arm times the reference's own CPU build on a bounded sample."""
import json
import os
import subprocess
import sys

import pytest

import helpers as H

sys.path.insert(0, H.ROOT)
import bench  # noqa: E402


CANNED = {
    "segments_total": 600_000_000, "kernel_seconds_max": 0.051, "e2e_seconds_max": 0.11, "segments_rank0": 600_000_000.0,
    "kernel_seconds_rank0": 0.051, "config": {"workload": "Coral2_P1 weak-scaled (canned)", "scale": 1.0},
    "clocks": {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 9},
    "h2d_bytes_per_step": 1426062498, "d2h_bytes_per_step": 1307394677, "gpu_launches": 44, "traffic": None,
    "whole_cycle": {"host_staged": {"cycle_init_ms": 383.0, "cycle_tracking_ms": 37.0, "cycle_finalize_ms": 0.04, "ms_per_cycle": 420.04,
                                    "segments_per_s_whole_cycle": 4.7e8},
                    "resident": {"cycles": 20, "cycle_init_kernel_ms_rank0": 0.628, "ms_per_cycle": 19.5}},
    "tracking_ms_per_step_rank0": {"boundary_particles_sent": 0, "cuda_events_on_kernel_stream": 17.0, "host_clock_around_call": 17.1},
    "balance_check": {"gains": 1, "losses": 1, "conserved": True, "last_row": []},
}


def _line(monkeypatch, capsys, evidence):
    from quicksilver_b200 import driver
    monkeypatch.setattr(driver, "run_benchmark", lambda *a, **k: json.loads(json.dumps(CANNED)))
    monkeypatch.setattr(bench, "kernel_evidence", evidence)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "3", "--warmup", "3", "--cpu-baseline", "0", "--extras", "0"])
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    assert bench.main() == 0
    return json.loads(capsys.readouterr().out.strip().splitlines()[-1])


def test_b200_line_carries_the_contract_keys(monkeypatch, capsys):
    evidence = {"dram_bytes_per_launch": 6.0e9, "thread_instructions_per_segment": 700.0, "kernel_hash": "feedfacefeedface"}
    line = _line(monkeypatch, capsys, lambda workload, scale: ("feedfacefeedface", evidence, "canned"))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "whole_cycle"):
        assert key in line, key
    assert line["unit"] == "segments/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] == pytest.approx(600_000_000 / 0.051)
    assert line["e2e"]["value"] == pytest.approx(600_000_000 / 0.11) and line["e2e"]["h2d_bytes_per_step"] > 0
    roof = line["roofline"]
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and roof["frac"] == pytest.approx(roof["achieved"] / roof["peak"])
    # evidence captured from the loaded kernel: the real traffic, the DRAM fraction and the lane-issue ceiling beside the yardstick
    assert roof["traffic"] == 6.0e9 and roof["dram_frac"] == pytest.approx(6.0e9 / 17.0e-3 / 1e9 / roof["peak"])
    assert roof["issue_frac"] == pytest.approx(700.0 * 2.0e8 / (148 * 4 * 32 * 1965e6 * 17.0e-3))
    init = line["whole_cycle"]["resident"]["cycle_init_roofline"]
    assert init["kernel"] == "cycle_init_kernel" and 0.5 < init["frac"] < 1.0


def test_profiler_evidence_of_another_kernel_is_refused(monkeypatch, capsys, tmp_path):
    """profiles/dram_traffic.json names the kernel it was captured from; a library built from different kernel sources must not
    inherit its numbers"""
    from quicksilver_b200 import _capi
    loaded = _capi.lib().qsb_kernel_hash().decode()
    assert len(loaded) == 16 and loaded != "unknown"
    prof = tmp_path / "profiles"
    prof.mkdir()
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    (prof / "dram_traffic.json").write_text(json.dumps({"Coral2_P1": {"kernel_hash": "0123456789abcdef", "dram_bytes_per_launch": 1}}))
    assert bench.kernel_evidence("Coral2_P1", 1.0)[1] is None
    (prof / "dram_traffic.json").write_text(json.dumps({"Coral2_P1": {"kernel_hash": loaded, "dram_bytes_per_launch": 7}}))
    assert bench.kernel_evidence("Coral2_P1", 1.0)[1]["dram_bytes_per_launch"] == 7
    assert bench.kernel_evidence("Coral2_P1", 0.5)[1] is None
    line = _line(monkeypatch, capsys, lambda workload, scale: (loaded, None, "refused"))
    assert line["roofline"]["traffic"] is None and "issue_frac" not in line["roofline"] and line["roofline"]["kernel_hash"] == loaded


@pytest.mark.skipif(not os.path.exists(H.REF_QS), reason="oracle/_ref/qs not built")
def test_reference_arm_times_the_reference_binary():
    out = subprocess.run([sys.executable, os.path.join(H.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                          "--workload", "Coral2_P2", "--scale", "0.03"], stdout=subprocess.PIPE, text=True, check=True, timeout=600).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["value"] > 1e5 and line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    # the arm runs the B200 arm's own per-GPU problem (same_config), on a bounded number of cycles, and says so
    cfg = line["config"]
    assert cfg["same_config_as_b200_arm_at_n1"] is True and cfg["scale"] == 0.03 and cfg["cells_per_gpu"] == 14 ** 3
    assert 1 <= cfg["cycles_timed"] <= 2 and line["cpu_baseline"]["host"]["usable_threads"] >= 1
    assert line["literal_deck_size"]["value"] > 1e5
