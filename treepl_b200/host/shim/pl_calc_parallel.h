// Drop-in shim: a header with the reference's file name (blackrim/treePL src/pl_calc_parallel.h).  Put this
// directory IN FRONT of the reference's own source directory on the include path (-I<repo>/treepl_b200/host/shim
// -I<reference>/src) and every `#include "pl_calc_parallel.h"` of the reference — optimize_tnc.cpp:17,
// optimize_nlopt.cpp:5, siman_calc_par.h, utils.h, main.cpp — resolves to the B200 engine's host mirror.
// No reference source file is edited.
#ifndef PL_CALC_PARALLEL_H_
#define PL_CALC_PARALLEL_H_

#include <map>
#include <string>
#include <vector>

#include "../pl_calc_b200.h"

using namespace std;      // the reference header does this, and its includers rely on it (pl_calc_parallel.h:15)

typedef pl_calc_b200 pl_calc_parallel;

#endif /* PL_CALC_PARALLEL_H_ */
