// Wrapper that makes dexopt's MLP / penalty / RNG operators available to the
// oracle recorders without copying them: the body is lines 22-375 of
// /root/reference/dexopt/src/ops.h (batch, dot4add, acos, relu,
// add_random_normal, add_random_uniform, dropout2, and the collision operators
// collision_project_2 / collision_project / collision_axes), extracted by the Makefile
// at build time into oracle/_ref/gen/dexopt_ops_lifted.inc.  The collision operators
// compile against shim/tractor_collision/collision_decl.h: they are registered (names,
// argument layouts, derivative rules) but cannot answer a query -- Bullet is absent.
// The rest of that file (:377-568, constraint operators) needs real Eigen and is not lifted.
#pragma once
#include <array>
#include <random>
#include <tractor/core/operator.h>
#include <tractor/core/recorder.h>
#include <tractor/core/var.h>
#include <tractor/geometry/vector3.h>
#include "shim/tractor_collision/collision_decl.h"
namespace tractor {
#include "dexopt_ops_lifted.inc"
}
