/* all.cu -- single translation unit of libglpb200 (kernels are shared by the
 * solver, the kernel-level entry points and the branch-and-bound driver) */
#include "solver.cu"
#include "kapi.cu"
#include "mip.cu"
#include "bnbpool.cu"
