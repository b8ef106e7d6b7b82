/* see adolc.h in this directory (oracle stub, test infrastructure only) */
#include <adolc/adolc.h>
