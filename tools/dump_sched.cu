// dump_sched.cu — prints the fused-kernel schedules (host only; debugging aid / used by tests/test_schedule.py)
#include <cstdio>
#include <cstring>
#include "../nerf-3dtalker-code_b200/csrc/hn_mlp_sched.h"
using namespace hn;
int main(int argc, char** argv) {
    const HostSchedules& hs = host_schedules();
    const bool bwd = argc > 1 && !strcmp(argv[1], "bwd");
    const int nu = bwd ? hs.bwd.n_ops : hs.fwd.n_ops, ne = bwd ? hs.bwd.n_epis : kFwdEpis;
    const MmaOp* mma = bwd ? hs.bwd.mma : hs.fwd.mma;
    const EpiOp* epi = bwd ? hs.bwd.epi : hs.fwd.epi;
    const PackOp* pk = bwd ? hs.bwd_pack : hs.fwd_pack;
    printf("units %d epis %d\n", nu, ne);
    for (int u = 0; u < nu; ++u)
        printf("U %d unit %d nkb %d a_blk %d n %d col %d q %d first %d commit %d wait_src %d wait_empty %d | w %d T %d l5 %d row0 %d col0 %d vr %d vc %d\n",
               u, mma[u].unit, mma[u].nkb, mma[u].a_blk, mma[u].n8 * 8, mma[u].tmem_col8 * 8, mma[u].q, mma[u].first, mma[u].commit, mma[u].wait_src, mma[u].wait_empty,
               pk[mma[u].unit].w_idx, pk[mma[u].unit].transposed, pk[mma[u].unit].l5_hidden, pk[mma[u].unit].row0, pk[mma[u].unit].col0, pk[mma[u].unit].valid_r, pk[mma[u].unit].valid_c);
    for (int e = 0; e < ne; ++e)
        printf("E %d q %d col %d width %d kind %d dst_blk %d ready %d density %d bias_off %d col0 %d save_blk %d mask_word %d\n",
               e, epi[e].q, epi[e].tmem_col8 * 8, epi[e].width32 * 32, epi[e].kind, epi[e].dst_blk, epi[e].ready_idx, epi[e].density,
               epi[e].bias_off, epi[e].col0, epi[e].save_blk, epi[e].mask_word);
    return 0;
}
