/* mip.cpp -- CPU ORACLE (test infrastructure only; see glpo.h). placeholder */
#include "glpo.h"
namespace glpo {
int intopt(Prob &P, const IOCP &parm, long *n_nodes) { (void)P; (void)parm; if (n_nodes) *n_nodes = 0; return GLP_EFAIL; }
}
