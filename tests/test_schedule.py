"""CPU-only protocol check of the fused kernels' static schedules (csrc/hn_mlp_sched.cu): replays the MMA-issuer and
epilogue programs against the mbarrier semantics and asserts (1) no deadlock over several tiles, (2) every barrier has
as many arrivals as waits per tile (otherwise phase parities drift), (3) the in-place rule: an epilogue never overwrites
activation blocks that an already-issued-but-uncommitted MMA of the same layer still reads."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dump():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = os.path.join(ROOT, "build", "dump_sched")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run([nvcc, "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe, os.path.join(ROOT, "tools", "dump_sched.cu"),
                    os.path.join(ROOT, "nerf-3dtalker-code_b200", "csrc", "hn_mlp_sched.cu")], check=True, capture_output=True)

    def run(which):
        out = subprocess.run([exe] + ([which] if which != "fwd" else []), check=True, capture_output=True, text=True).stdout
        mma, epi = [], []
        for line in out.splitlines():
            kv = dict(re.findall(r"(\w+) (-?\d+)", line.split("|")[0]))
            if line.startswith("U "):
                mma.append({k: int(v) for k, v in kv.items()})
            elif line.startswith("E "):
                epi.append({k: int(v) for k, v in kv.items()})
        return mma, epi
    return run


@pytest.mark.parametrize("which", ["fwd", "bwd"])
def test_schedule_protocol(dump, which):
    mma, epi = dump(which)
    assert mma and epi
    # (2) arrivals == waits per barrier per tile
    for c in range(3):
        arrivals = sum(1 for e in epi if e["ready"] == c)
        waits = sum(1 for m in mma if m["wait_src"] == 1 + c)
        assert arrivals == waits, f"a_ready[{c}]: {arrivals} arrivals vs {waits} waits"
    for q in range(4):
        assert sum(1 for m in mma if m["commit"] and m["q"] == q) == sum(1 for e in epi if e["q"] == q), f"acc_full[{q}]"
        assert sum(1 for m in mma if m["wait_empty"] and m["q"] == q) == sum(1 for e in epi if e["q"] == q), f"acc_empty[{q}]"
    assert all(m["wait_empty"] for m in mma if m["first"]), "an accumulator is overwritten without waiting for its release"

    # (1) replay three tiles
    n_tiles = 3
    ready = [0, 0, 0]; ready_used = [0, 0, 0]
    full = [0] * 4; full_used = [0] * 4
    empty = [0] * 4; empty_used = [0] * 4            # releases / consumed-by-MMA (first use is free)
    pm = pe = 0
    total_m, total_e = len(mma) * n_tiles, len(epi) * n_tiles
    while pm < total_m or pe < total_e:
        progressed = False
        if pm < total_m:
            m = mma[pm % len(mma)]
            ok = True
            if m["wait_src"] in (1, 2, 3):
                c = m["wait_src"] - 1
                ok = ready[c] > ready_used[c]
            if ok and m["wait_empty"]:
                ok = empty_used[m["q"]] == 0 or empty[m["q"]] >= empty_used[m["q"]]
            if ok:
                if m["wait_src"] in (1, 2, 3):
                    ready_used[m["wait_src"] - 1] += 1
                if m["wait_empty"]:
                    empty_used[m["q"]] += 1
                if m["commit"]:
                    full[m["q"]] += 1
                pm += 1
                progressed = True
        if pe < total_e:
            e = epi[pe % len(epi)]
            if full[e["q"]] > full_used[e["q"]]:
                full_used[e["q"]] += 1
                empty[e["q"]] += 1
                if e["ready"] != 255:
                    ready[e["ready"]] += 1
                pe += 1
                progressed = True
        assert progressed, f"deadlock at MMA op {pm} / epilogue op {pe}"

    # (3) in-place rule
    for q in range(4):
        commits = [i for i, m in enumerate(mma) if m["commit"] and m["q"] == q]
        eps = [e for e in epi if e["q"] == q]
        for ci, e in zip(commits, eps):
            if e["ready"] == 255 or e["kind"] in (2, 6):
                continue
            blocks = {e["dst_blk"], e["dst_blk"] + 1} if e["width"] > 64 else {e["dst_blk"]}
            fenced = False
            for m in mma[ci + 1:]:
                if m["wait_src"] == 1 + e["ready"]:
                    fenced = True                     # the next layer's first reader waits for this epilogue
                touched = {m["a_blk"] + k for k in range(m.get("nkb", 1))}
                if touched & blocks and m["a_blk"] != 6:
                    assert fenced, f"MMA unit {m['unit']} reads blocks {touched & blocks} that epilogue (q={q}) is overwriting"
                    break
