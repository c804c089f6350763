// hn_mlp_sched.cu — host-side generation of the fused-kernel schedules (see hn_mlp_sched.h).
#include <cassert>
#include <cstdlib>
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
// Forward chain (A operand in tensor memory, alternating TMEM halves): table generation, see hn_mlp_sched.h.
struct FwdLayer {
    int w_idx; int smem_kb; int n_kb; int widths[3]; int n_chunks; EpiKind kind; int bias_off; int save_blk; int mask_word; bool l5_hidden; int density; bool to_tmem;
    bool transposed = false;        // data-gradient chain: units hold W^T (rows = layer inputs, columns = layer outputs)
};

// HN_SPLIT_TAIL=0 restores the unsplit phase A (A/B measurements); read once, before the tables are built
static bool split_tail_enabled() {
    static const bool on = [] { const char* e = getenv("HN_SPLIT_TAIL"); return !(e && atoi(e) == 0); }();
    return on;
}

// Generates the stage / pack / epilogue tables of a GEMM chain whose activations live in tensor memory.  `smem_kb` leading
// K blocks of a layer come from shared memory: the forward's positional-encoding block (1, awaited once per tile) or the
// data-gradient chain's streamed dL/dfeat blocks (4, each awaited).
void build_tmem_chain(const std::vector<FwdLayer>& layers, bool streamed_input, FwdTables* T, PackOp* pack_out, int n_units_expected, int n_epis_expected) {
    std::vector<StageOp> stages;
    std::vector<PackOp> pack;
    std::vector<EpiOp2> epi;
    std::vector<int> X;                                                // TMEM column of every input K block (even-tile columns)
    int chunk0 = 0;
    auto pack_unit = [&](const FwdLayer& L, int j, int k) {            // weight unit (chunk j, K block k); k < 0: empty unit
        PackOp p{};
        p.w_idx = (int8_t)L.w_idx; p.transposed = (int8_t)L.transposed;
        if (k < 0) { p.valid_r = 0; p.valid_c = 0; pack.push_back(p); return; }
        if (L.transposed) {                                            // unit(r,c) = W[(row0 + c) * ld + col0 + r]
            p.row0 = (int16_t)(64 * k); p.valid_c = 64;
            p.col0 = (int16_t)(128 * j); p.valid_r = (int16_t)L.widths[j]; p.l5_hidden = (int8_t)L.l5_hidden;
            pack.push_back(p);
            return;
        }
        const bool is_pe = k < L.smem_kb;
        const int kb = k - L.smem_kb;
        p.row0 = (int16_t)(128 * j); p.valid_r = (int16_t)L.widths[j];
        if (is_pe) { p.col0 = 0; p.valid_c = HN_PE; p.l5_hidden = 0; }
        else { p.col0 = (int16_t)(64 * kb); p.valid_c = 64; p.l5_hidden = (int8_t)L.l5_hidden; }
        pack.push_back(p);
    };
    for (size_t li = 0; li < layers.size(); ++li) {
        const FwdLayer& L = layers[li];
        const int B = (li & 1) ? 256 : 0;                              // accumulator half of this layer (even tiles)
        const int n_k = L.smem_kb + L.n_kb;
        auto a_src = [&](int k) -> uint16_t { return (k < L.smem_kb) ? (uint16_t)(kSrcSmem | (streamed_input ? k : 0)) : (uint16_t)X[k - L.smem_kb]; };
        auto wait_of = [&](int k) -> uint8_t {                        // whoever meets an input slot first waits for it
            if (k < L.smem_kb) return (streamed_input || li == 0) ? 4 : 0;
            const int kb = k - L.smem_kb;
            return (kb % 2 == 0) ? (uint8_t)(1 + kb / 2) : 0;
        };
        // phase A: chunks 0 and 1 together, one stage per K block ...
        // ... except the LAST TWO K blocks (split tail): they are the ones the previous layer's last epilogue releases, so everything
        // issued after them sits on the layer-to-layer critical path (epilogue of layer l's last chunk -> these MMAs -> epilogue of
        // layer l+1's first chunk).  Chunk 0 takes both of them alone (N = 128) and commits; chunk 1 follows (N = 128 / 64) and
        // commits: chunk 0's epilogue starts half a tail (8 MMAs) earlier and chunk 1's remainder runs underneath it.
        const bool split_tail = split_tail_enabled() && n_k >= 3 && !(a_src(n_k - 1) & kSrcSmem) && !(a_src(n_k - 2) & kSrcSmem);
        const int n_full = split_tail ? n_k - 2 : n_k;
        for (int k = 0; k < n_full; ++k) {
            StageOp s{};
            s.a_src0 = a_src(k); s.a_src1 = kSrcNone;
            s.acc_col = (uint16_t)B;
            s.n8 = (uint8_t)((L.widths[0] + L.widths[1]) / 8);
            s.first = (uint8_t)(k == 0);
            s.commit = (uint8_t)((!split_tail && k == n_k - 1) ? 2 : 0);
            s.chunk = (uint8_t)chunk0;
            s.wait_src = wait_of(k);
            stages.push_back(s);
            pack_unit(L, 0, k); pack_unit(L, 1, k);
        }
        if (split_tail) {
            const uint8_t w0 = wait_of(n_k - 2), w1 = wait_of(n_k - 1);
            assert(!(w0 && w1));                                       // one input barrier per stage
            for (int j = 0; j < 2; ++j) {
                StageOp s{};
                s.a_src0 = a_src(n_k - 2); s.a_src1 = a_src(n_k - 1);
                s.acc_col = (uint16_t)(B + 128 * j);
                s.n8 = (uint8_t)(L.widths[j] / 8);
                s.first = 0;
                s.commit = 1;
                s.chunk = (uint8_t)(chunk0 + j);
                s.wait_src = (uint8_t)(j == 0 ? (w0 ? w0 : w1) : 0);
                stages.push_back(s);
                pack_unit(L, j, n_k - 2); pack_unit(L, j, n_k - 1);
            }
        }
        // phase B: chunk 2 alone over [B, B+128), two K blocks per stage (the PE block, read from shared memory, gets a stage
        // of its own next to an empty unit: the issuer handles one kind of A operand per stage)
        if (L.n_chunks == 3) {
            assert(!(streamed_input && L.smem_kb));                    // streamed blocks are consumed once, in phase A only
            std::vector<std::pair<int, int>> sts;
            int k = 0;
            if (L.smem_kb) { sts.push_back({0, -1}); k = 1; }
            for (; k < n_k; k += 2) sts.push_back({k, k + 1 < n_k ? k + 1 : -1});
            for (size_t i = 0; i < sts.size(); ++i) {
                StageOp s{};
                s.a_src0 = a_src(sts[i].first); s.a_src1 = sts[i].second >= 0 ? a_src(sts[i].second) : kSrcNone;
                s.acc_col = (uint16_t)B;
                s.n8 = 16;
                s.first = (uint8_t)(i == 0);
                s.commit = (uint8_t)(i + 1 == sts.size() ? 1 : 0);
                s.chunk = (uint8_t)(chunk0 + 2);
                s.wait_p = (uint8_t)(i == 0);
                stages.push_back(s);
                pack_unit(L, 2, sts[i].first); pack_unit(L, 2, sts[i].second);
            }
        }
        // epilogue ops and where their packed outputs go
        std::vector<int> Y;
        for (int j = 0; j < L.n_chunks; ++j) {
            const int width = L.widths[j];
            int acc = B + (j == 1 ? 128 : 0), out = -1, wait_next = 0;
            if (L.to_tmem) {
                if (L.n_chunks == 3) { out = j == 0 ? B + 192 : (j == 1 ? B + 128 : B); wait_next = (j == 0); }
                else { out = j == 0 ? B + 192 : B + 128; }              // 192-wide layers: [B+192, B+256) is outside the accumulator
            }
            EpiOp2 e{};
            e.acc_col = (uint16_t)acc;
            e.out_col = out < 0 ? kNoCol : (uint16_t)out;
            e.width32 = (uint8_t)(width / 32);
            e.kind = L.kind;
            e.ready_idx = L.to_tmem ? (uint8_t)j : 255;
            e.flags = (uint8_t)((L.density ? (j == L.n_chunks - 1 ? 2 : 1) : 0) | (wait_next ? 4 : 0) | ((L.n_chunks == 3 && j == 0) ? 8 : 0));
            e.bias_off = (uint16_t)(L.bias_off + 128 * j);
            e.col0 = (uint16_t)(128 * j);
            e.save_blk = L.save_blk < 0 ? 0xFFFF : (uint16_t)(L.save_blk + 2 * j);
            e.mask_word = L.mask_word < 0 ? 0xFFFF : (uint16_t)(L.mask_word + 4 * j);
            epi.push_back(e);
            if (out >= 0) for (int h = 0; h < width / 64; ++h) Y.push_back(out + 32 * h);
        }
        chunk0 += L.n_chunks;
        X = Y;
    }
    assert((int)stages.size() * 2 == n_units_expected && (int)pack.size() == n_units_expected && (int)epi.size() == n_epis_expected);
    assert((int)stages.size() <= kFwdStages && (int)epi.size() <= kFwdEpis);
    memcpy(pack_out, pack.data(), sizeof(PackOp) * pack.size());
    memcpy(T->stage, stages.data(), sizeof(StageOp) * stages.size());
    memcpy(T->epi, epi.data(), sizeof(EpiOp2) * epi.size());
    T->n_stages = (int)stages.size();
    T->n_epis = (int)epi.size();
    // Across tiles the same no-barrier argument must hold between the LAST layer of a tile and the FIRST of the next: their
    // accumulators must sit in different halves.  An odd layer count needs the halves swapped on odd tiles, an even one does not.
    T->tile_flip = (int)(layers.size() & 1);
    for (const EpiOp2& e : epi) if (e.ready_idx != 255) T->n_ready[e.ready_idx]++;
}

void build_fwd(HostSchedules* hs) {
    // ---- layer list: NetWorks/models.py:69-82
    std::vector<FwdLayer> layers;
    for (int l = 0; l < 8; ++l)
        layers.push_back({l, (l == 0 || l == 5) ? 1 : 0, l == 0 ? 0 : 6, {128, 128, 128}, 3, EPI_HIDDEN, l * HN_HIDDEN, HN_SLOT_H0 + 6 * l, 12 * l, l == 5, l == 7 ? 1 : 0, true});
    // (RGB_layer_0 is folded into RGB_layer_1's weights and bias: see hn_mlp_sched.h)
    layers.push_back({W_R1, 0, 6, {128, 64, 0}, 2, EPI_HIDDEN, HN_BIAS_OFF_R1, HN_SLOT_X, 96, false, 0, true});
    layers.push_back({W_R2, 0, 3, {128, 128, 0}, 2, EPI_FEAT, HN_BIAS_OFF_R2, -1, -1, false, 0, false});
    build_tmem_chain(layers, false, &hs->fwd, hs->fwd_pack, kFwdUnits, kFwdEpis);
    // the PE block is free once FeaExt_module_5's last chunk has completed: the next tile's PE is produced there
    hs->fwd.pe_after_epi = 17;
}

// data-gradient chain without dL/dPE (autograd of NetWorks/models.py:62-87 in reverse): dX = dZ * W, then the ReLU mask of the
// layer below; RGB_layer_0's input gradient also receives the density head's rank-1 term
void build_bwdt(HostSchedules* hs) {
    std::vector<FwdLayer> layers;
    FwdLayer r2{W_R2, 4, 0, {128, 64, 0}, 2, EPI_GRAD_MASK, 0, HN_GSLOT_R1, 96, false, 0, true}; r2.transposed = true;
    layers.push_back(r2);
    // (W_R1[:, :384] W_R0)^T: straight to dL/d(FeaExt_module_7 pre-activation), which also receives the density head's rank-1 term
    FwdLayer r1{W_R1, 0, 3, {128, 128, 128}, 3, EPI_GRAD_DENSITY, 0, HN_GSLOT_Z0 + 6 * 7, 12 * 7, false, 0, true}; r1.transposed = true;
    layers.push_back(r1);
    for (int l = 7; l >= 1; --l) {
        // FeaExt_module_1^T closes the chain: its output (dZ of FeaExt_module_0) is only saved, no GEMM reads it
        FwdLayer s{l, 0, 6, {128, 128, 128}, 3, EPI_GRAD_MASK, 0, HN_GSLOT_Z0 + 6 * (l - 1), 12 * (l - 1), l == 5, 0, l > 1}; s.transposed = true;
        layers.push_back(s);
    }
    build_tmem_chain(layers, true, &hs->bwdt, hs->bwdt_pack, kBwdTUnits, kBwdTEpis);
    hs->bwdt.pe_after_epi = -1;
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
    build_fwd(hs);
    build_bwdt(hs);
    for (int with_pe = 1; with_pe >= 0; --with_pe) {
        // ---- data-gradient chain (reverse order); dL/dPE accumulated at FeaExt_module_5 and _0
        Builder b; b.n_acc = 3; b.backward = true;
        ChainStep r2{}; r2.w_idx = W_R2; r2.n_kb = 4; r2.n_out = HN_RGB1; r2.kind = EPI_GRAD_MASK; r2.save_blk = HN_GSLOT_R1;
        r2.mask_word = 96; b.step(r2, true);
        ChainStep r1{}; r1.w_idx = W_R1; r1.n_kb = 3; r1.n_out = HN_HIDDEN; r1.kind = EPI_GRAD_DENSITY;       // fused (W_R1 W_R0)^T
        r1.save_blk = HN_GSLOT_Z0 + 6 * 7; r1.mask_word = 12 * 7; b.step(r1, false);
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
