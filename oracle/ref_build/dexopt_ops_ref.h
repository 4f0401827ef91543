// Wrapper that makes dexopt's MLP / penalty / RNG operators available to the
// oracle recorders without copying them: the body is lines 22-204 of
// /root/reference/dexopt/src/ops.h (batch, dot4add, acos, relu,
// add_random_normal, add_random_uniform, dropout2), extracted by the Makefile
// at build time into oracle/_ref/gen/dexopt_ops_lifted.inc.  The rest of that
// file (collision_* ops, :208-568) needs real Eigen + Bullet and is not lifted.
#pragma once
#include <array>
#include <random>
#include <tractor/core/operator.h>
#include <tractor/core/recorder.h>
#include <tractor/core/var.h>
#include <tractor/geometry/vector3.h>
namespace tractor {
#include "dexopt_ops_lifted.inc"
}
