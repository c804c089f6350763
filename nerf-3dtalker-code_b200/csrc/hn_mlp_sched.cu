// hn_mlp_sched.cu — host-side generation of the fused-kernel schedules (see hn_mlp_sched.h).
#include <cassert>
#include <cstring>
#include <mutex>
#include <vector>
#include "hn_mlp_sched.h"
#include "../../include/headnerf_b200.h"

namespace hn {
namespace {

struct ChainStep {
    int w_idx;            // weight tensor
    bool pe_src;          // forward: a PE K-block precedes the activation K-blocks
    int n_kb;             // K-blocks read from the activation/gradient buffer
    int n_out;            // output width routed to 128-column chunks
    EpiKind kind;
    int bias_off;         // forward only
    int save_blk;         // first block of the save slot, -1 = none
    int mask_word;        // first mask word, -1 = none
    int w_col0;           // column of the first activation K-block in the weight (forward) / of output col 0 (backward)
    bool l5_hidden;       // w_col0 is relative to the hidden block of FeaExt_module_5
    int density;          // forward: 1 = accumulate density head in this step's epilogue
    bool pe_out;          // backward: an extra 64-wide chunk accumulates dL/dPE
    bool pe_first;        // backward: that chunk overwrites (first contribution of the tile)
    bool pe_only;         // backward: step has no regular output (FeaExt_module_0)
    bool no_consumer;     // no later GEMM reads this step's output from shared memory: do not signal a_ready
};

struct Builder {
    std::vector<PackOp> pack;
    std::vector<MmaOp> mma;
    std::vector<EpiOp> epi;
    int qc = 0;
    int n_acc;            // rotating accumulator chunks (4 forward, 3 backward)
    bool backward;

    void step(const ChainStep& s, bool first_step) {
        const int nch = s.pe_only ? 0 : (s.n_out + 127) / 128;
        std::vector<std::vector<int>> groups;
        if (s.pe_src) groups.push_back({kPeBlk});
        for (int c = 0; 2 * c < s.n_kb; ++c) {
            std::vector<int> g{2 * c};
            if (2 * c + 1 < s.n_kb) g.push_back(2 * c + 1);
            groups.push_back(g);
        }
        const int G = (int)groups.size();
        const int n_agroups = G - (s.pe_src ? 1 : 0);
        std::vector<int> q_of(nch);
        for (int j = 0; j < nch; ++j) q_of[j] = (qc + j) % n_acc;
        // in-place rule: the chunk that overwrites the K-blocks of the LAST group must complete last
        std::vector<int> last_order;
        for (int j = 0; j < nch; ++j) if (j != n_agroups - 1) last_order.push_back(j);
        if (n_agroups - 1 < nch && n_agroups >= 1) last_order.push_back(n_agroups - 1);
        std::vector<int> nat_order;
        for (int j = 0; j < nch; ++j) nat_order.push_back(j);

        for (int g = 0; g < G; ++g) {
            const bool is_pe_group = s.pe_src && g == 0;
            // chunk list of this K-group: the dL/dPE chunk (-1) goes FIRST so that no MMA still reads the
            // group's K-blocks after the regular chunk that overwrites them has been committed
            std::vector<int> chunks;
            if (s.pe_out) chunks.push_back(-1);
            for (int j : ((g == G - 1) ? last_order : nat_order)) chunks.push_back(j);
            for (size_t ci = 0; ci < chunks.size(); ++ci) {
                const int j = chunks[ci];
                const bool pe_chunk = (j < 0);
                for (size_t kk = 0; kk < groups[g].size(); ++kk) {
                    const int blk = groups[g][kk];
                    const int width = pe_chunk ? 64 : std::min(128, s.n_out - 128 * j);
                    MmaOp m{};
                    m.a_blk = (uint8_t)blk;
                    m.n8 = (uint8_t)(width / 8);
                    m.tmem_col8 = (uint8_t)((pe_chunk ? 384 : q_of[j] * 128) / 8);
                    m.q = (uint8_t)(pe_chunk ? 3 : q_of[j]);
                    const bool first_k = (g == 0 && kk == 0);
                    m.first = pe_chunk ? (uint8_t)(s.pe_first && first_k) : (uint8_t)first_k;
                    m.wait_empty = pe_chunk ? (uint8_t)(s.pe_first && first_k) : (uint8_t)first_k;
                    m.commit = (uint8_t)(g == G - 1 && kk + 1 == groups[g].size() && (!pe_chunk || s.pe_only));
                    // source readiness is waited for once, on the first unit that touches the group
                    if (ci == 0 && kk == 0) {
                        if (is_pe_group) m.wait_src = first_step ? 4 : 0;
                        else if (first_step && backward) m.wait_src = (g == 0) ? 5 : 0;
                        else m.wait_src = (uint8_t)(1 + blk / 2);
                    }
                    mma.push_back(m);

                    PackOp p{};
                    p.w_idx = (int8_t)s.w_idx;
                    p.l5_hidden = (int8_t)s.l5_hidden;
                    if (!backward) {
                        p.transposed = 0;
                        p.row0 = (int16_t)(128 * j);
                        p.valid_r = (int16_t)width;
                        if (blk == kPeBlk) { p.col0 = 0; p.valid_c = HN_PE; p.l5_hidden = 0; }
                        else { p.col0 = (int16_t)(s.w_col0 + 64 * blk); p.valid_c = 64; }
                    } else {
                        p.transposed = 1;                       // unit(r,c) = W[(row0+c)*ld + col0 + r]
                        p.row0 = (int16_t)(64 * blk);           // contraction index = layer output channel
                        p.valid_c = 64;
                        if (pe_chunk) { p.col0 = 0; p.valid_r = HN_PE; p.l5_hidden = 0; }
                        else { p.col0 = (int16_t)(s.w_col0 + 128 * j); p.valid_r = (int16_t)width; }
                    }
                    pack.push_back(p);
                }
            }
        }
        // epilogue ops in completion order
        for (int j : last_order) {
            EpiOp e{};
            e.q = (uint8_t)q_of[j];
            e.tmem_col8 = (uint8_t)(q_of[j] * 128 / 8);
            e.width32 = (uint8_t)(std::min(128, s.n_out - 128 * j) / 32);
            e.kind = s.kind;
            e.dst_blk = (uint8_t)(2 * j);
            e.ready_idx = (s.kind == EPI_FEAT || s.no_consumer) ? 255 : (uint8_t)j;
            e.density = (uint8_t)(s.density ? (j == nch - 1 ? 2 : 1) : 0);
            e.bias_off = (uint16_t)(s.bias_off + 128 * j);
            e.col0 = (uint16_t)(128 * j);
            e.save_blk = s.save_blk < 0 ? 0xFFFF : (uint16_t)(s.save_blk + 2 * j);
            e.mask_word = s.mask_word < 0 ? 0xFFFF : (uint16_t)(s.mask_word + 4 * j);
            epi.push_back(e);
        }
        if (s.pe_only) {
            EpiOp e{};
            e.q = 3; e.tmem_col8 = 384 / 8; e.width32 = 2; e.kind = EPI_GRAD_PE; e.ready_idx = 255;
            e.save_blk = 0xFFFF; e.mask_word = 0xFFFF;
            epi.push_back(e);
        }
        qc += nch;
    }
};


// ------------------------------------------------------------------------------------------------------------------
// Forward chain, second generation (A operand in tensor memory): slot allocator + table generation.
struct SlotSim {
    // state of each 64-column TMEM slot: 0 busy; 1 free, last touched by MMAs (ordered by issue order);
    // 2 free once the epilogue of chunk rel_n has loaded its accumulator
    int kind[8], rel_n[8];
    bool ok_acc(int s, int n) const { return kind[s] == 1 || (kind[s] == 2 && rel_n[s] <= n - 2); }
    bool ok_out(int s, int n) const { return kind[s] == 1 || (kind[s] == 2 && rel_n[s] <= n); }
};

struct Fwd2Layer { int w_idx; bool pe; int n_kb; int widths[3]; int n_chunks; EpiKind kind; int bias_off; int save_blk; int mask_word; bool l5_hidden; int density; bool to_tmem; };

struct Chunk2 { int acc_slot; int acc_slots; int out_slot; int wait_prev; };

// returns false if `fixed` (when given) violates a hazard rule from the initial state `sim`
bool run_slots(const std::vector<Fwd2Layer>& layers, SlotSim& sim, int n0, std::vector<Chunk2>& table, bool fixed) {
    std::vector<int> X;
    int n = n0;
    size_t idx = 0;
    for (const Fwd2Layer& L : layers) {
        std::vector<int> Y;
        for (int j = 0; j < L.n_chunks; ++j, ++n, ++idx) {
            const int need = L.widths[j] == 128 ? 2 : 1;
            Chunk2 c{};
            if (fixed) c = table[idx];
            else {
                c.acc_slots = need; c.acc_slot = -1;
                if (need == 2) {
                    int best = -1, best_cost = 99;
                    for (int a = 0; a < 8; a += 2)
                        if (sim.ok_acc(a, n) && sim.ok_acc(a + 1, n)) {
                            const int cost = (sim.kind[a] != 1) + (sim.kind[a + 1] != 1);
                            if (cost < best_cost) { best = a; best_cost = cost; }
                        }
                    c.acc_slot = best;
                } else {
                    int best = -1, best_cost = 99;
                    for (int a = 0; a < 8; ++a)
                        if (sim.ok_acc(a, n)) {
                            const int cost = 2 * (sim.kind[a ^ 1] != 0) + (sim.kind[a] != 1);     // keep free pairs intact
                            if (cost < best_cost) { best = a; best_cost = cost; }
                        }
                    c.acc_slot = best;
                }
                if (c.acc_slot < 0) return false;
            }
            for (int k = 0; k < need; ++k) { if (!sim.ok_acc(c.acc_slot + k, n)) return false; sim.kind[c.acc_slot + k] = 0; }
            if (!L.to_tmem) c.out_slot = -1;
            else if (!fixed) {
                c.out_slot = -1;
                for (int a = 0; a < 8 && c.out_slot < 0; ++a)            // a lone free slot (its pair partner is not free)
                    if (sim.ok_out(a, n) && !sim.ok_out(a ^ 1, n)) c.out_slot = a;
                if (c.out_slot < 0) c.out_slot = c.acc_slot;            // in place over the first accumulator slot
            }
            if (c.out_slot >= 0 && !(c.out_slot >= c.acc_slot && c.out_slot < c.acc_slot + need)) {
                if (!sim.ok_out(c.out_slot, n)) return false;
                const int wp = (sim.kind[c.out_slot] == 2 && sim.rel_n[c.out_slot] == n - 1) ? 1 : 0;
                if (fixed && wp && !c.wait_prev) return false;
                if (!fixed) c.wait_prev = wp;
                sim.kind[c.out_slot] = 0;
            }
            for (int k = 0; k < need; ++k)
                if (c.acc_slot + k != c.out_slot) { sim.kind[c.acc_slot + k] = 2; sim.rel_n[c.acc_slot + k] = n; }
            if (c.out_slot >= 0) Y.push_back(c.out_slot);
            if (!fixed) table.push_back(c);
        }
        for (int s : X) sim.kind[s] = 1;                               // this layer's inputs: free once its MMAs are issued
        X = Y;
    }
    for (int s : X) sim.kind[s] = 1;
    return true;
}

void build_fwd2(HostSchedules* hs) {
    // ---- layer list: NetWorks/models.py:69-82
    std::vector<Fwd2Layer> layers;
    for (int l = 0; l < 8; ++l)
        layers.push_back({l, l == 0 || l == 5, l == 0 ? 0 : 6, {128, 128, 128}, 3, EPI_HIDDEN, l * HN_HIDDEN, HN_SLOT_H0 + 6 * l, 12 * l, l == 5, l == 7 ? 1 : 0, true});
    layers.push_back({W_R0, false, 6, {128, 128, 128}, 3, EPI_LINEAR, HN_BIAS_OFF_R0, HN_SLOT_R0, -1, false, 0, true});
    layers.push_back({W_R1, false, 6, {128, 64, 0}, 2, EPI_HIDDEN, HN_BIAS_OFF_R1, HN_SLOT_X, 96, false, 0, true});
    layers.push_back({W_R2, false, 3, {128, 128, 0}, 2, EPI_FEAT, HN_BIAS_OFF_R2, -1, -1, false, 0, false});

    // ---- TMEM slots: derive the table from an empty TMEM, then prove it is also valid when a tile starts from the
    // state the previous tile (same table) leaves behind
    std::vector<Chunk2> table;
    SlotSim sim{};
    for (int s = 0; s < 8; ++s) { sim.kind[s] = 1; sim.rel_n[s] = 0; }
    bool ok = run_slots(layers, sim, 0, table, false);
    assert(ok && (int)table.size() == kFwdEpis);
    for (int s = 0; s < 8; ++s) { assert(sim.kind[s] != 0); sim.rel_n[s] -= kFwdEpis; }
    ok = run_slots(layers, sim, 0, table, true);
    assert(ok && "forward TMEM slot table is not periodic");
    (void)ok;

    // ---- tables
    std::vector<MmaOp2> mma;
    std::vector<PackOp> pack;
    std::vector<EpiOp2> epi;
    std::vector<int> X;                                                // TMEM column of every input K block
    size_t idx = 0;
    for (size_t li = 0; li < layers.size(); ++li) {
        const Fwd2Layer& L = layers[li];
        std::vector<int> Y;
        for (int j = 0; j < L.n_chunks; ++j, ++idx) {
            const Chunk2& c = table[idx];
            const int width = L.widths[j];
            const int acc_col = c.acc_slot * 64;
            const int n_kb = (L.pe ? 1 : 0) + L.n_kb;
            for (int k = 0; k < n_kb; ++k) {
                const bool is_pe = L.pe && k == 0;
                const int kb = k - (L.pe ? 1 : 0);
                MmaOp2 m{};
                m.a_src = is_pe ? (uint16_t)(kSrcSmem | 0) : (uint16_t)X[kb];
                m.acc_col = (uint16_t)acc_col;
                m.n8 = (uint8_t)(width / 8);
                m.first = (uint8_t)(k == 0);
                m.commit = (uint8_t)(k == n_kb - 1);
                m.wait_src = 0;
                if (j == 0) {                                           // the first chunk of a layer meets every input block first
                    if (is_pe) m.wait_src = (li == 0) ? 4 : 0;
                    else if (kb % 2 == 0) m.wait_src = (uint8_t)(1 + kb / 2);
                }
                mma.push_back(m);
                PackOp p{};
                p.w_idx = (int8_t)L.w_idx; p.transposed = 0;
                p.row0 = (int16_t)(128 * j); p.valid_r = (int16_t)width;
                if (is_pe) { p.col0 = 0; p.valid_c = HN_PE; p.l5_hidden = 0; }
                else { p.col0 = (int16_t)(64 * kb); p.valid_c = 64; p.l5_hidden = (int8_t)L.l5_hidden; }
                pack.push_back(p);
            }
            EpiOp2 e{};
            e.acc_col = (uint16_t)acc_col;
            e.out_col = c.out_slot < 0 ? kNoCol : (uint16_t)(c.out_slot * 64);
            e.width32 = (uint8_t)(width / 32);
            e.kind = L.kind;
            e.ready_idx = L.to_tmem ? (uint8_t)j : 255;
            e.density = (uint8_t)(L.density ? (j == L.n_chunks - 1 ? 2 : 1) : 0);
            e.wait_prev = (uint8_t)c.wait_prev;
            e.bias_off = (uint16_t)(L.bias_off + 128 * j);
            e.col0 = (uint16_t)(128 * j);
            e.save_blk = L.save_blk < 0 ? 0xFFFF : (uint16_t)(L.save_blk + 2 * j);
            e.mask_word = L.mask_word < 0 ? 0xFFFF : (uint16_t)(L.mask_word + 4 * j);
            epi.push_back(e);
            if (c.out_slot >= 0) for (int h = 0; h < width / 64; ++h) Y.push_back(c.out_slot * 64 + 32 * h);
        }
        X = Y;
    }
    assert((int)mma.size() == kFwdUnits && (int)epi.size() == kFwdEpis);
    memcpy(hs->fwd_pack, pack.data(), sizeof(PackOp) * kFwdUnits);
    memcpy(hs->fwd.mma, mma.data(), sizeof(MmaOp2) * kFwdUnits);
    memcpy(hs->fwd.epi, epi.data(), sizeof(EpiOp2) * kFwdEpis);
    hs->fwd.n_ops = kFwdUnits;
    // the PE block is free once FeaExt_module_5's last chunk has completed: the next tile's PE is produced there
    hs->fwd.pe_after_epi = 17;
    for (const EpiOp2& e : epi) if (e.ready_idx != 255) hs->fwd.n_ready[e.ready_idx]++;
}

// Two consecutive weight units that continue the same accumulator chunk over the next K block become one
// MMA op (one barrier wait / commit per 32 KiB of weights instead of per 16 KiB).
std::vector<MmaOp> merge_k_pairs(const std::vector<MmaOp>& in) {
    std::vector<MmaOp> out;
    for (size_t i = 0; i < in.size(); ++i) {
        MmaOp m = in[i];
        m.nkb = 1;
        if (i + 1 < in.size()) {
            const MmaOp& n = in[i + 1];
            if (!m.commit && n.q == m.q && n.tmem_col8 == m.tmem_col8 && n.n8 == m.n8 && n.a_blk == m.a_blk + 1 && m.a_blk != kPeBlk &&
                !n.first && !n.wait_src && !n.wait_empty && n.unit == m.unit + 1) {
                m.nkb = 2;
                m.commit = n.commit;
                ++i;
            }
        }
        out.push_back(m);
    }
    return out;
}

HostSchedules* build() {
    auto* hs = new HostSchedules();
    memset(hs, 0, sizeof(*hs));
    build_fwd2(hs);
    for (int with_pe = 1; with_pe >= 0; --with_pe) {
        // ---- data-gradient chain (reverse order); dL/dPE accumulated at FeaExt_module_5 and _0
        Builder b; b.n_acc = 3; b.backward = true;
        ChainStep r2{}; r2.w_idx = W_R2; r2.n_kb = 4; r2.n_out = HN_RGB1; r2.kind = EPI_GRAD_MASK; r2.save_blk = HN_GSLOT_R1;
        r2.mask_word = 96; b.step(r2, true);
        ChainStep r1{}; r1.w_idx = W_R1; r1.n_kb = 3; r1.n_out = HN_HIDDEN; r1.kind = EPI_GRAD_LINEAR; r1.save_blk = HN_GSLOT_R0;
        r1.mask_word = -1; b.step(r1, false);
        ChainStep r0{}; r0.w_idx = W_R0; r0.n_kb = 6; r0.n_out = HN_HIDDEN; r0.kind = EPI_GRAD_DENSITY;
        r0.save_blk = HN_GSLOT_Z0 + 6 * 7; r0.mask_word = 12 * 7; b.step(r0, false);
        for (int l = 7; l >= 1; --l) {
            ChainStep s{};
            s.w_idx = l; s.n_kb = 6; s.n_out = HN_HIDDEN; s.kind = EPI_GRAD_MASK;
            s.save_blk = HN_GSLOT_Z0 + 6 * (l - 1); s.mask_word = 12 * (l - 1);
            s.l5_hidden = (l == 5); s.pe_out = (l == 5) && with_pe; s.pe_first = s.pe_out;
            s.no_consumer = (l == 1) && !with_pe;
            b.step(s, false);
        }
        if (with_pe) {
            ChainStep l0{}; l0.w_idx = W_L0; l0.n_kb = 6; l0.n_out = 0; l0.pe_out = true; l0.pe_only = true; l0.kind = EPI_GRAD_PE;
            l0.save_blk = -1; l0.mask_word = -1; b.step(l0, false);
        }
        assert((int)b.mma.size() <= kBwdUnitsMax && (int)b.epi.size() <= kBwdEpisMax);
        BwdTables& t = with_pe ? hs->bwd : hs->bwd_nope;
        t.n_epis = (int)b.epi.size();
        if (with_pe) hs->n_bwd_pack_units = (int)b.mma.size();
        if (with_pe) {
            memcpy(hs->bwd_pack, b.pack.data(), sizeof(PackOp) * b.pack.size());
            for (size_t u = 0; u < b.mma.size(); ++u) b.mma[u].unit = (uint16_t)u;
        } else {
            for (size_t u = 0; u < b.mma.size(); ++u) {
                int found = -1;
                for (int v = 0; v < hs->n_bwd_pack_units; ++v)
                    if (!memcmp(&hs->bwd_pack[v], &b.pack[u], sizeof(PackOp))) { found = v; break; }
                assert(found >= 0);
                b.mma[u].unit = (uint16_t)found;
            }
        }
        std::vector<MmaOp> merged = merge_k_pairs(b.mma);
        memcpy(t.mma, merged.data(), sizeof(MmaOp) * merged.size());
        t.n_ops = (int)merged.size();
        memcpy(t.epi, b.epi.data(), sizeof(EpiOp) * b.epi.size());
        for (const EpiOp& e : b.epi) { if (e.ready_idx != 255) t.n_ready[e.ready_idx]++; t.n_empty[e.q]++; }
    }
    return hs;
}

}  // namespace

const HostSchedules& host_schedules() {
    static HostSchedules* hs = build();
    return *hs;
}

}  // namespace hn
