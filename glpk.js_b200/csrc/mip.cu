/* mip.cu -- branch-and-bound driver (placeholder until the ios restatement lands) */
#include "glpb_internal.cuh"
extern "C" int glpb_intopt(glpb_prob *P, const glpb_iocp *parm) { (void)parm; if (!P) return GLPB_EINVAL; return GLPB_ESTATE; }
extern "C" int glpb_get_mip(glpb_prob *P, int *mip_stat, double *mip_obj, double *mipx, long *nodes)
{
    if (!P) return GLPB_EINVAL;
    if (mip_stat) *mip_stat = P->mip_stat;
    if (mip_obj) *mip_obj = P->mip_obj;
    if (mipx) for (int k = 0; k < P->m + P->n; k++) mipx[k] = P->h_mipx[k];
    if (nodes) *nodes = P->mip_nodes;
    return 0;
}
