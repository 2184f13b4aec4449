"""CPU checks of bench.py's contract plumbing: byte accounting, sharding bounds, the clock sampler's parsing and the
reference arm's single JSON line (the GPU arm is exercised on the GPU box)."""
import json
import os
import stat
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_uvd_byte_accounting():
    n, r = 100_000_000, 10
    per, step, survey = bench.uvd_bytes(n, r, "fused")
    assert step == 4 * n * (7 * r + 11) == per[1] + per[9] + per[13]            # the three sweeps of the fused call
    assert survey == 4 * n * (9 * r + 12) == per[1] + per[2] + per[3] + per[4] + per[5]   # two separate calls
    _, step2, _ = bench.uvd_bytes(n, r, "separate")
    assert step2 == survey
    assert per[1] + per[7] + per[8] + per[4] + per[5] == survey                 # two-sweep update + d pass + apply


def test_chunks_tile_the_vector_on_256_row_boundaries():
    for n in (1, 255, 1021, 100_000_000):
        for world in (1, 2, 3, 8):
            b = [bench.chunk_of(n, world, k) for k in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            for (lo0, hi0), (lo1, hi1) in zip(b, b[1:]):
                assert hi0 == lo1 and lo1 % 256 == 0 or lo1 == n


def test_clock_sampler_keeps_only_samples_of_the_timed_region(tmp_path, monkeypatch):
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/usr/bin/env python3\n"
                    "import time, datetime\n"
                    "time.sleep(0.2)\n"
                    "i = 0\n"
                    "while True:\n"
                    "    ts = datetime.datetime.now().strftime('%Y/%m/%d %H:%M:%S.%f')[:-3]\n"
                    "    mhz = 1000 if i < 3 else 1900\n"
                    "    print(f'{ts}, {mhz}, 1965, 700.5, 0x4, Not Active, Not Active, Not Active, Active', flush=True)\n"
                    "    i += 1\n"
                    "    time.sleep(0.05)\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", f"{tmp_path}:{os.environ['PATH']}")
    c = bench.ClockSampler(0)
    c.start()                       # returns once the first sample is on disk: start-up stays outside the timed region
    time.sleep(0.3)                 # "warm-up": the 1000 MHz samples
    c.mark()
    time.sleep(0.25)
    out = c.stop()
    assert out is not None and out["samples"] >= 3
    assert out["sm_mhz"] == 1900.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]


def test_reference_arm_prints_exactly_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "uvd",
                        "--steps", "1", "--warmup", "0", "--n", "200000"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
