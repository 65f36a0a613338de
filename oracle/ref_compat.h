// ref_compat.h -- force-included when compiling the reference: covkernel.cpp:494-527 call unqualified
// isnan/isinf, which g++ 13 only provides in namespace std.  TEST INFRASTRUCTURE ONLY.
#ifdef __cplusplus
#include <cmath>
using std::isnan;
using std::isinf;
#endif
