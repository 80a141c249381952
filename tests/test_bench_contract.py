"""CPU suite: the reference arm of bench.py runs without a GPU and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().split("\n")[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True
    assert line["metric"] == "domain pixel*GN-evaluations/s" and line["unit"] == "pixel*evaluations/s"
    assert line["value"] > 0 and line["cpu_baseline"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "c1" in line["config"]["workload"]


def test_subset_boxes_follow_manager_arithmetic():
    sys.path.insert(0, ROOT)
    import bench
    boxes = bench.subset_boxes(64, 8128, 64)
    assert len(boxes) == 4096
    w = {b[2] - b[0] + 1 for b in boxes} | {b[3] - b[1] + 1 for b in boxes}
    assert w == {125}                       # (8064 / 64 - 1) / 2 = 62 -> 2 * 62 + 1
    assert boxes[0][:2] == (65, 65) and boxes[1][0] == boxes[0][0]  # centre int(.5 + 64 + 62.5) = 127; sector = i * vs + j, j (y) fastest
    assert all(b[0] >= 64 and b[2] <= 8128 for b in boxes)


def test_upload_band_covers_domain_plus_halo_and_stays_inside_the_image():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.upload_band(0, 100, 8192, 2) == (0, 100 + 1 + 64 + 32)
    b = bench.upload_band(4000, 5000, 8192, 2)
    assert b == (4000 - 96, 5001 + 96)
    assert bench.upload_band(8000, 8191, 8192, 4) == (8000 - 192, 8192)
    # every level's valid rows (two fewer per level at each cut) still cover the domain at that level
    lo, hi = bench.upload_band(4000, 5000, 8192, 4)
    for lv in range(1, 5):
        lo, hi = (lo + 3) // 2, (hi - 3) // 2 + 1
        assert lo <= (4000 >> lv) - 2 and hi >= (5000 >> lv) + 3


class _FakeStagingEngine:
    """The staging rules of dic_stage_next_pair / dic_advance_pair / dic_correlate*_async / _wait
    (include/dic_b200.h), without a device: at most two pairs staged, advance needs one, a wait needs an enqueue."""

    def __init__(self):
        self.staged = []      # pair ids staged and not yet current
        self.current = None   # pair id the next solve reads
        self.in_flight = None
        self.next_id = 0
        self.solved = []
        self.max_staged = 0

    def stage(self):
        assert len(self.staged) < 2, "a third pair was staged without an advance"
        self.staged.append(self.next_id)
        self.next_id += 1
        self.max_staged = max(self.max_staged, len(self.staged))

    def advancePair(self):
        assert self.staged, "advance without a staged pair"
        assert self.in_flight is None, "advance while a solve still reads the current pair's records"
        self.current = self.staged.pop(0)

    def enqueue(self):
        assert self.current is not None and self.in_flight is None
        self.in_flight = self.current

    def wait(self):
        assert self.in_flight is not None, "wait without an enqueued solve"
        self.solved.append(self.in_flight)
        self.in_flight = None


def test_e2e_loop_stages_two_pairs_ahead_and_solves_every_pair_once():
    """bench.Run.e2e_loop on a fake engine: n steps stage n pairs, solve pairs 0..n-1 in order, keep two pairs
    staged ahead once the pipeline is full, and leave nothing staged or in flight."""
    sys.path.insert(0, ROOT)
    import bench
    for n in (1, 2, 3, 7, 32):
        fake = _FakeStagingEngine()
        run = bench.Run.__new__(bench.Run)
        run.eng = fake
        run.stage_pair = fake.stage
        run.enqueue_step = fake.enqueue
        run.wait_step = fake.wait
        run.step_work = lambda: 1.0
        assert run.e2e_loop(n) == float(n)
        assert fake.solved == list(range(n)) and fake.next_id == n
        assert fake.staged == [] and fake.in_flight is None
        assert fake.max_staged == min(n, 2)
