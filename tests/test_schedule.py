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
            kv = dict(re.findall(r"(\w+) (-?\d+)", line.replace("|", " ")))
            if line.startswith("U "):
                mma.append({k: int(v) for k, v in kv.items()})
            elif line.startswith("E "):
                epi.append({k: int(v) for k, v in kv.items()})
        return mma, epi
    return run


def _replay_fwd(mma, epi, n_tiles, late):
    """Replays the forward chain (A operand in tensor memory) against the two ordering facts the kernel provides
    (MMAs execute in issue order; chunk n's MMAs wait for the accumulator loads of chunk n-2; epilogue stores of chunk n
    follow the loads of every chunk <= n) with the epilogues running as EARLY or as LATE as those facts allow, tagging
    every 32-column TMEM block with what was last written there.  Asserts that every MMA reads the K block it expects
    and every epilogue drains its own, unclobbered accumulator."""
    tm = {}                                           # 32-column block -> tag
    n_chunks = len(epi)
    loads_done = stores_done = 0                      # global epilogue progress (chunk counters)
    commits = 0
    ready = [0, 0, 0]; ready_used = [0, 0, 0]

    def epi_load(n):
        t, e = divmod(n, n_chunks)
        op = epi[e]
        for c in range(op["acc_col"], op["acc_col"] + op["width"], 32):
            assert tm.get(c) == ("acc", n), f"epilogue {e} (tile {t}) finds {tm.get(c)} at column {c}"

    def epi_store(n):
        t, e = divmod(n, n_chunks)
        op = epi[e]
        if op["out_col"] >= 0:
            for h in range(op["width"] // 64):
                tm[op["out_col"] + 32 * h] = ("act", n, h)
        if op["ready"] != 255:
            ready[op["ready"]] += 1

    def run_epilogues(until_loads, until_stores):
        nonlocal loads_done, stores_done
        while loads_done < until_loads or stores_done < until_stores:
            if stores_done < loads_done:              # a warp stores chunk n before it loads chunk n+1
                epi_store(stores_done); stores_done += 1
            else:
                assert loads_done < commits, "epilogue would wait for an accumulator that is never committed"
                epi_load(loads_done); loads_done += 1

    for t in range(n_tiles):
        # producer chunk of every a_src column is whatever tag sits there when the layer's first chunk reads it; later
        # chunks of the same layer must see the very same tags
        seen = {}
        for u, m in enumerate(mma):
            if m["wait_src"] in (1, 2, 3):
                c = m["wait_src"] - 1
                ready_used[c] += 1
                # the MMA issuer blocks until that epilogue has stored: force exactly as much epilogue progress as needed
                while ready[c] < ready_used[c]:
                    run_epilogues(min(stores_done + 1, commits), stores_done + 1)
            chunk = t * n_chunks + m["chunk"]
            if m["first"] and chunk >= 2:
                run_epilogues(max(chunk - 1, loads_done), stores_done)  # accumulator loads of chunk n-2 done
            if not late:
                run_epilogues(commits, commits)
            if not m["smem"]:
                tag = tm.get(m["a_src"])
                assert tag is not None and tag[0] == "act", f"MMA unit {u} (tile {t}) reads {tag} at column {m['a_src']}"
                seen.setdefault((layer_first_of(mma, u), m["a_src"]), tag)
                assert seen[(layer_first_of(mma, u), m["a_src"])] == tag, f"MMA unit {u}: K block at column {m['a_src']} changed under the layer"
            for c in range(m["acc_col"], m["acc_col"] + m["n"], 32):
                tm[c] = ("acc", chunk)
            if m["commit"]:
                assert chunk == commits, "chunks must complete in epilogue order"
                commits += 1
    run_epilogues(commits, commits)


def layer_first_of(mma, u):
    """index of the first unit of the layer that unit u belongs to (layers start where a_ready[0] / pe_ready is awaited)"""
    while not (mma[u]["wait_src"] in (1, 4) and mma[u]["first"]):
        u -= 1
    return u


@pytest.mark.parametrize("late", [False, True])
def test_forward_tmem_schedule(dump, late):
    mma, epi = dump("fwd")
    assert len(mma) == 168 and len(epi) == 31
    for c in range(3):
        arrivals = sum(1 for e in epi if e["ready"] == c)
        waits = sum(1 for m in mma if m["wait_src"] == 1 + c)
        assert arrivals == waits, f"a_ready[{c}]: {arrivals} arrivals vs {waits} waits"
    assert sum(m["commit"] for m in mma) == len(epi) and sum(m["first"] for m in mma) == len(epi)
    # every layer's K blocks come from the previous layer's chunks in order (chunk j -> K blocks 2j, 2j+1)
    _replay_fwd(mma, epi, 3, late)
    # weight units follow the (layer, chunk, K block) stream
    for u, m in enumerate(mma):
        assert m["vr"] == m["n"] and m["row0"] % 128 == 0
    firsts = [m["chunk"] for m in mma if m["first"]]
    assert firsts == sorted(firsts), "chunks must start in order (the n-2 accumulator rule relies on it)"


@pytest.mark.parametrize("which", ["bwd"])
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
