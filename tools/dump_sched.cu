// dump_sched.cu — prints the fused-kernel schedules (host only; debugging aid / used by tests/test_schedule.py)
#include <cstdio>
#include <cstring>
#include "../nerf-3dtalker-code_b200/csrc/hn_mlp_sched.h"
using namespace hn;
static void print_pack(const PackOp& p) {
    printf(" | w %d T %d l5 %d row0 %d col0 %d vr %d vc %d\n", p.w_idx, p.transposed, p.l5_hidden, p.row0, p.col0, p.valid_r, p.valid_c);
}
int main(int argc, char** argv) {
    const HostSchedules& hs = host_schedules();
    const bool bwd = argc > 1 && !strcmp(argv[1], "bwd");
    const bool bwdt = argc > 1 && !strcmp(argv[1], "bwdt");
    if (!bwd) {
        // forward chain / tensor-memory data-gradient chain: A operand in tensor memory (StageOp / EpiOp2)
        const FwdTables& T = bwdt ? hs.bwdt : hs.fwd;
        const PackOp* pk = bwdt ? hs.bwdt_pack : hs.fwd_pack;
        printf("stages %d epis %d pe_after %d tile_flip %d\n", T.n_stages, T.n_epis, T.pe_after_epi, T.tile_flip);
        for (int u = 0; u < T.n_stages; ++u) {
            const StageOp& m = T.stage[u];
            printf("S %d chunk %d smem0 %d a0 %d smem1 %d a1 %d acc_col %d n %d first %d commit %d wait_src %d wait_p %d", u, m.chunk, (m.a_src0 & kSrcSmem) ? 1 : 0,
                   m.a_src0 & 0x7FFF, (m.a_src1 != kSrcNone && (m.a_src1 & kSrcSmem)) ? 1 : 0, m.a_src1 == kSrcNone ? -1 : (m.a_src1 & 0x7FFF), m.acc_col, m.n8 * 8,
                   m.first, m.commit, m.wait_src, m.wait_p);
            print_pack(pk[2 * u]);
            printf("P %d", 2 * u + 1); print_pack(pk[2 * u + 1]);
        }
        for (int e = 0; e < T.n_epis; ++e) {
            const EpiOp2& o = T.epi[e];
            printf("E %d acc_col %d out_col %d width %d kind %d ready %d density %d wait_next %d signal_p %d bias_off %d col0 %d save_blk %d mask_word %d\n", e, o.acc_col,
                   o.out_col == kNoCol ? -1 : (int)o.out_col, o.width32 * 32, o.kind, o.ready_idx, o.flags & 3, (o.flags >> 2) & 1, (o.flags >> 3) & 1, o.bias_off, o.col0, o.save_blk, o.mask_word);
        }
        return 0;
    }
    const int nu = hs.bwd.n_ops, ne = hs.bwd.n_epis;
    const MmaOp* mma = hs.bwd.mma;
    const EpiOp* epi = hs.bwd.epi;
    const PackOp* pk = hs.bwd_pack;
    printf("units %d epis %d\n", nu, ne);
    for (int u = 0; u < nu; ++u) {
        printf("U %d unit %d nkb %d a_blk %d n %d col %d q %d first %d commit %d wait_src %d wait_empty %d",
               u, mma[u].unit, mma[u].nkb, mma[u].a_blk, mma[u].n8 * 8, mma[u].tmem_col8 * 8, mma[u].q, mma[u].first, mma[u].commit, mma[u].wait_src, mma[u].wait_empty);
        print_pack(pk[mma[u].unit]);
    }
    for (int e = 0; e < ne; ++e)
        printf("E %d q %d col %d width %d kind %d dst_blk %d ready %d density %d bias_off %d col0 %d save_blk %d mask_word %d\n",
               e, epi[e].q, epi[e].tmem_col8 * 8, epi[e].width32 * 32, epi[e].kind, epi[e].dst_blk, epi[e].ready_idx, epi[e].density,
               epi[e].bias_off, epi[e].col0, epi[e].save_blk, epi[e].mask_word);
    return 0;
}
