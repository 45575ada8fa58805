"""CPU checks of the measurement helpers that have host-only logic (scripts/zgemm_trace.py: the CARC_ZGEMM_TRACE log summary)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_zgemm_trace_summary(tmp_path):
    log = tmp_path / "trace.log"
    log.write_text(
        "some unrelated stderr line\n"
        "zgemm 4096 4096 4096 1 0 0 1 16.0000\n"
        "zgemm 4096 4096 4096 1 0 0 1 16.5000\n"
        "zgemm 8 64 64 4096 1 0 1 0.7000\n"
        "zgemm truncated line\n")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "zgemm_trace.py"), str(log)], capture_output=True,
                         text=True, check=True).stdout.splitlines()
    assert out[0].startswith("total 33.2 ms in 3 products, 2 shapes")
    rows = [line.split() for line in out[2:]]
    assert rows[0][:4] == ["4096", "4096", "4096", "1"] and rows[0][7] == "2"       # sorted by time, two calls summed
    tflops = float(rows[0][-1])
    assert abs(tflops - 8 * 4096 ** 3 * 2 / 32.5e-3 / 1e12) < 0.01
    assert rows[1][:4] == ["8", "64", "64", "4096"]
