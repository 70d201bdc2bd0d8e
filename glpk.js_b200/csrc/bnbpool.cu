/* bnbpool.cu -- C ABI of the batched, device-resident branch-and-bound
 * (nodeengine.cuh: one node per CTA; bnbpool.cuh: the tree on the host).
 * glpb_intopt uses it for every problem the node engine takes (m <= 64,
 * n <= 1024: the reference's test/gap.lpt, test/todd.lpt and the multi-
 * dimensional knapsack of BASELINE.json configs[4]); larger problems keep
 * the one-LP-at-a-time driver of mip.cu.
 */
static double glpb_now_ms() { return now_ms(); }
#include "bnbpool.cuh"

void glpb_bnb_free(glpb_bnb *T) { delete T; }

static int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

/* true when glpb_intopt should take the batched path */
static bool bnb_batched(const glpb_prob *P, const glpb_iocp &parm)
{
    const char *mode = getenv("GLPB_BNB");
    if (mode && !strcmp(mode, "serial")) return false;
    if (!(parm.br_tech >= GLP_BR_FFV && parm.br_tech <= GLP_BR_DTH)) return false;
    return glpb_bnb::eligible(P);
}

extern "C" int glpb_bnb_begin(glpb_prob *P, const glpb_iocp *parm_, int batch, int slab_nodes)
{
    if (!P) return GLPB_EINVAL;
    glpb_iocp parm;
    if (parm_) parm = *parm_; else glpb_init_iocp(&parm);
    if (!(parm.br_tech >= GLP_BR_FFV && parm.br_tech <= GLP_BR_DTH)) {
        glpb_set_error("glp_intopt: br_tech = %d not supported on the device path", parm.br_tech);
        return GLPB_EINVAL;
    }
    if (!(parm.bt_tech >= GLP_BT_DFS && parm.bt_tech <= GLP_BT_BPH)) return GLPB_EINVAL;
    if (!(0.0 < parm.tol_int && parm.tol_int < 1.0) || !(0.0 < parm.tol_obj && parm.tol_obj < 1.0)) return GLPB_EINVAL;
    if (!glpb_bnb::eligible(P)) {
        glpb_set_error("batched branch-and-bound takes m <= %d, n <= %d", NE_MAXM, NE_MAXN);
        return GLPB_EINVAL;
    }
    P->mip_stat = GLP_UNDEF;
    P->mip_obj = 0.0;
    int rc = mip_check(P);
    if (rc) return rc;
    CK(cudaSetDevice(P->device));
    if (P->bnb) { glpb_bnb_free(P->bnb); P->bnb = nullptr; }
    glpb_bnb *T = new glpb_bnb();
    T->tm_beg = now_ms();
    if (batch <= 0) batch = env_int("GLPB_BNB_BATCH", 16 * (P->sm_count > 0 ? P->sm_count : 148));
    if (slab_nodes <= 0) slab_nodes = env_int("GLPB_BNB_SLAB", 262144);
    rc = T->init(P, parm, batch, slab_nodes);
    if (rc) { delete T; return rc; }
    P->bnb = T;
    return 0;
}

/* one round = one launch of k_bnb_nodes over up to `batch` open nodes.
   Returns 1 after a round, 0 if the local pool was empty, GLP_E* when stopped
   (time / node limit, mip gap, failure).  *done = nodes of this round. */
extern "C" int glpb_bnb_round(glpb_prob *P, long max_tasks, long *done)
{
    if (!P || !P->bnb) return GLPB_ESTATE;
    CK(cudaSetDevice(P->device));
    int rc = P->bnb->round(max_tasks, done);
    P->mip_nodes = P->bnb->solved;
    return rc;
}

extern "C" int glpb_bnb_open_count(glpb_prob *P) { return (!P || !P->bnb) ? GLPB_ESTATE : P->bnb->open_count(); }

extern "C" int glpb_bnb_get_incumbent(glpb_prob *P, int *has_solution, double *obj)
{
    if (!P || !P->bnb) return GLPB_ESTATE;
    if (has_solution) *has_solution = P->bnb->have_sol;
    if (obj) *obj = P->bnb->have_cut ? P->bnb->mip_obj : (P->dir == GLP_MIN ? +DBL_MAX : -DBL_MAX);
    return 0;
}

extern "C" int glpb_bnb_set_cutoff(glpb_prob *P, double obj)
{
    if (!P || !P->bnb) return GLPB_ESTATE;
    P->bnb->set_cutoff(obj);
    return 0;
}

extern "C" int glpb_bnb_clear(glpb_prob *P)
{
    if (!P || !P->bnb) return GLPB_ESTATE;
    P->bnb->clear();
    return 0;
}

extern "C" long glpb_bnb_record_bytes(glpb_prob *P) { return (!P || !P->bnb) ? GLPB_ESTATE : (long)P->bnb->record_bytes(); }

/* node migration: dev_buf is DEVICE memory of the handle's GPU (e.g. the
   storage of a torch tensor that NCCL then gathers); the node payload is
   copied device-to-device and never touches the host */
extern "C" int glpb_bnb_export_nodes(glpb_prob *P, int max_count, void *dev_buf, int *count)
{
    if (!P || !P->bnb || !dev_buf || !count) return GLPB_ESTATE;
    CK(cudaSetDevice(P->device));
    return P->bnb->export_nodes(max_count, (unsigned char *)dev_buf, count);
}

extern "C" int glpb_bnb_import_nodes(glpb_prob *P, const void *dev_buf, int count)
{
    if (!P || !P->bnb || (!dev_buf && count > 0)) return GLPB_ESTATE;
    CK(cudaSetDevice(P->device));
    return P->bnb->import_nodes((const unsigned char *)dev_buf, count);
}

/* out[0] node LPs solved, [1] nodes processed, [2] rounds (= launches),
   [3] dual simplex iterations, [4] basis inversions, [5] open nodes,
   [6] bytes of shared memory per CTA, [7] matrix resident in shared memory,
   [8..13] SM cycles summed over the nodes (thread 0 of every CTA): preprocessing, basis
   inversions, fresh bbar/cbar, simplex iterations, integrality + branching, node state I/O */
extern "C" int glpb_bnb_stats(glpb_prob *P, long *out, int count)
{
    if (!P || !P->bnb || !out) return GLPB_ESTATE;
    glpb_bnb &T = *P->bnb;
    long v[14] = {T.solved, T.tasks_done, T.rounds, T.iters, T.refacs, (long)T.open_count(), (long)T.smem_bytes, (long)T.np.a_in_smem,
                  (long)T.cyc[0], (long)T.cyc[1], (long)T.cyc[2], (long)T.cyc[3], (long)T.cyc[4], (long)T.cyc[5]};
    for (int i = 0; i < count && i < 14; i++) out[i] = v[i];
    return 0;
}

/* the status mapping of solve_mip (lib/glpapi09.js:82-112); ret is the code
   the search ended with (0 = tree exhausted on every rank) */
extern "C" int glpb_bnb_end(glpb_prob *P, int ret)
{
    if (!P || !P->bnb) return GLPB_ESTATE;
    glpb_bnb &T = *P->bnb;
    P->mip_nodes = T.solved;
    if (T.have_sol) {
        P->mip_stat = (ret == 0) ? GLP_OPT : GLP_FEAS;
        P->mip_obj = T.mip_obj;
        P->h_mipx = T.mipx;
    } else if (T.have_cut) { P->mip_stat = GLP_UNDEF; P->mip_obj = T.mip_obj; }   /* optimum found on another rank */
    else P->mip_stat = (ret == 0) ? GLP_NOFEAS : GLP_UNDEF;
    glpb_bnb_free(P->bnb);
    P->bnb = nullptr;
    return ret;
}

static int bnb_intopt_batched(glpb_prob *P, const glpb_iocp *parm)
{
    int rc = glpb_bnb_begin(P, parm, 0, 0);
    if (rc) return rc;
    int ret;
    for (;;) {
        ret = glpb_bnb_round(P, -1, nullptr);
        if (ret != 1) break;
    }
    return glpb_bnb_end(P, ret);
}
