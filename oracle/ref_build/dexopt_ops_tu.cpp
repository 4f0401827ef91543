// Instantiates dexopt's operators (see dexopt_ops_ref.h) exactly the way
// /root/reference/dexopt/src/ops.cpp does: define TRACTOR_IMPLEMENT_OPS, include.
#define TRACTOR_IMPLEMENT_OPS
#include "dexopt_ops_ref.h"
