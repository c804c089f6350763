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
        run.tile_flip = 1
        for line in out.splitlines():
            if line.startswith("stages ") and "tile_flip" in line:
                run.tile_flip = int(line.split("tile_flip")[1])
            kv = dict(re.findall(r"(\w+) (-?\d+)", line.replace("|", " ")))
            if line.startswith("U ") or line.startswith("S "):
                mma.append({k: int(v) for k, v in kv.items()})
            elif line.startswith("E "):
                epi.append({k: int(v) for k, v in kv.items()})
        return mma, epi
    return run


def _replay_fwd(stages, epi, n_tiles, late, tile_flip=1):
    """Replays the forward chain (A operand in tensor memory, alternating TMEM halves) against the ordering facts the kernel
    provides - MMAs execute in issue order; a stage with wait_src waits for that a_ready barrier, one with wait_p for the
    accumulator loads of chunk-2; the epilogue takes chunks in order, loads a chunk only after its commit, stores chunk 0 of
    a pair together with chunk 1 (after chunk 1's loads), any other chunk after its own loads - with the epilogue running as
    EARLY or as LATE as those facts allow.  Every 32-column TMEM block is tagged with what was last written there; every MMA
    must read the K block it expects and every epilogue must drain its own, unclobbered accumulator."""
    tm = {}
    n_chunks = len(epi)
    loads_done = stores_done = 0                      # chunks whose accumulator has been loaded / whose output has been stored
    commits = 0
    ready = [0, 0, 0]; ready_used = [0, 0, 0]

    def hx(n):
        return 256 if (n // n_chunks) & tile_flip else 0

    def do_load(n):
        op = epi[n % n_chunks]
        for c in range(op["acc_col"], op["acc_col"] + op["width"], 32):
            assert tm.get(c ^ hx(n)) == ("acc", n), f"epilogue of chunk {n} finds {tm.get(c ^ hx(n))} at column {c ^ hx(n)}"

    def do_store(n):
        op = epi[n % n_chunks]
        if op["out_col"] >= 0:
            for h in range(op["width"] // 64):
                tm[(op["out_col"] + 32 * h) ^ hx(n)] = ("act", n, h)
        if op["ready"] != 255:
            ready[op["ready"]] += 1

    def step_epilogue():
        """one more epilogue event, in program order; returns False if it has to wait for a commit"""
        nonlocal loads_done, stores_done
        n = loads_done
        pending_store = stores_done < loads_done      # a loaded chunk whose output is not stored yet
        if pending_store:
            m = stores_done
            if epi[m % n_chunks]["wait_next"] and loads_done == m + 1:
                pass                                  # held: stored after the NEXT chunk's loads
            else:
                do_store(m); stores_done += 1
                return True
        if n >= commits:
            return False
        do_load(n); loads_done += 1
        return True

    def run_until(cond):
        while not cond():
            assert step_epilogue(), "deadlock: the epilogue waits for a commit the MMA issuer cannot reach"

    seen = {}
    for t in range(n_tiles):
        base = t * n_chunks
        layer_key = None
        for u, m in enumerate(stages):
            if m["wait_src"] in (1, 2, 3):
                c = m["wait_src"] - 1
                ready_used[c] += 1
                run_until(lambda: ready[c] >= ready_used[c])
            if m["wait_p"]:
                run_until(lambda: loads_done > base + m["chunk"] - 2)
            if not late:
                while step_epilogue():
                    pass
            if m["first"] and m["wait_src"] in (1, 4):
                layer_key = (t, u)
            for a_src, smem in ((m["a0"], m["smem0"]), (m["a1"], m["smem1"])):
                if a_src < 0 or smem:
                    continue
                col = a_src ^ (256 if t & tile_flip else 0)
                tag = tm.get(col)
                assert tag is not None and tag[0] == "act", f"stage {u} (tile {t}) reads {tag} at column {col}"
                seen.setdefault((layer_key, a_src), tag)
                assert seen[(layer_key, a_src)] == tag, f"stage {u}: K block at column {col} changed under the layer"
            for c in range(m["acc_col"], m["acc_col"] + m["n"], 32):
                chunk = base + m["chunk"] + (1 if (m["n"] > 128 and c >= m["acc_col"] + 128) else 0)
                tm[c ^ (256 if t & tile_flip else 0)] = ("acc", chunk)
            for k in range(m["commit"]):
                assert base + m["chunk"] + k == commits, "chunks must complete in epilogue order"
                commits += 1
    while step_epilogue():
        pass
    assert loads_done == commits == n_tiles * n_chunks and stores_done == commits


@pytest.mark.parametrize("late", [False, True])
def test_forward_tmem_schedule(dump, late):
    stages, epi = dump("fwd")
    assert len(stages) == 76 and len(epi) == 28          # 10 GEMMs: RGB_layer_0 is folded into RGB_layer_1 (csrc/hn_mlp_sched.h)
    for c in range(3):
        arrivals = sum(1 for e in epi if e["ready"] == c)
        waits = sum(1 for m in stages if m["wait_src"] == 1 + c)
        assert arrivals == waits, f"a_ready[{c}]: {arrivals} arrivals vs {waits} waits"
    assert sum(m["commit"] for m in stages) == len(epi)
    # a chunk whose output is held for its neighbour must be followed by a chunk of the same layer that stores
    for i, e in enumerate(epi):
        if e["wait_next"]:
            assert i + 1 < len(epi) and not epi[i + 1]["wait_next"] and epi[i + 1]["ready"] == 1 and e["ready"] == 0
    assert dump.tile_flip == 0                      # even number of layers: the TMEM halves must NOT swap between tiles
    _replay_fwd(stages, epi, 4, late, tile_flip=0)
    with pytest.raises(AssertionError):             # ... and the replay does catch the cross-tile clobber if they did
        _replay_fwd(stages, epi, 4, True, tile_flip=1)


@pytest.mark.parametrize("late", [False, True])
def test_data_gradient_tmem_schedule(dump, late):
    """The data-gradient chain without dL/dPE runs on the same tensor-memory machinery as the forward chain."""
    stages, epi = dump("bwdt")
    assert len(stages) == 72 and len(epi) == 26
    for c in range(3):
        assert sum(1 for e in epi if e["ready"] == c) == sum(1 for m in stages if m["wait_src"] == 1 + c), f"a_ready[{c}]"
    assert sum(m["commit"] for m in stages) == len(epi)
    # the four dL/dfeat K blocks stream through a two-block shared-memory ring: each is awaited, in order, in phase A only
    smem = [m for m in stages if m["smem0"]]
    assert [m["a0"] for m in smem] == [0, 1, 2, 3] and all(m["wait_src"] == 4 and m["a1"] < 0 for m in smem)
    for i, e in enumerate(epi):
        if e["wait_next"]:
            assert i + 1 < len(epi) and not epi[i + 1]["wait_next"] and epi[i + 1]["ready"] == 1 and e["ready"] == 0
    assert dump.tile_flip == 1                      # nine layers: odd tiles swap the TMEM halves
    _replay_fwd(stages, epi, 4, late, tile_flip=1)
    with pytest.raises(AssertionError):             # ... and the replay does catch the cross-tile clobber if they did not
        _replay_fwd(stages, epi, 4, True, tile_flip=0)


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
