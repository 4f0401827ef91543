#!/bin/bash
# Regenerates the golden fixtures in this directory from the REFERENCE's own code.
# Needs /root/reference (this container only); run `make -C oracle/ref_build` first.
# Every array in these files was computed by the reference SimpleEngine
# (tractor/src/engines/simple.cpp:22-35) on programs recorded through the reference
# Var<> API and differentiated by the reference buildGradients() (gradients.cpp:22-407).
set -euo pipefail
cd "$(dirname "$0")/../.."
R=oracle/_ref/ref_tool
G=tests/golden
$R registry $G/ref_registry.txt
$R optests $G/optests_8d.trxb suffix=8d seed=7
$R optests $G/optests_d.trxb suffix=d seed=7
$R record $G/fkchain24.trxb problem=fkchain joints=24 tiles=2 gd_steps=10
$R record /tmp/dexsyn_small.trxb problem=dexsyn frames=3 hidden=4 tiles=2 extra_fixed=4 gd_steps=10
$R record /tmp/dexsyn_dropout_outer2.trxb problem=dexsyn frames=2 hidden=4 tiles=2 extra_fixed=2 dropout=1 outer=2 gd_steps=10
xz -9 -f -c /tmp/dexsyn_small.trxb > $G/dexsyn_small.trxb.xz
xz -9 -f -c /tmp/dexsyn_dropout_outer2.trxb > $G/dexsyn_dropout_outer2.trxb.xz
# values-only fixtures (programs=0) for the re-rolled rollout objective at BASELINE's shapes: the product
# records / assembles these problems itself (trx_dexsyn_build, trx_objective_create_rollout)
$R record $G/rollout_handarm_h64.trxb problem=dexsyn kind=handarm hidden=64 frames=4 tiles=2 gd_steps=10 programs=0 gd_keep=last keep_dr=0
$R record $G/rollout_hand8_linear.trxb problem=dexsyn kind=hand8 hidden=0 frames=5 tiles=2 gd_steps=10 programs=0 gd_keep=last keep_dr=0
$R record $G/rollout_chain30.trxb problem=dexsyn kind=chain joints=30 contacts=16 hidden=16 frames=3 tiles=1 programs=0 keep_dr=0
$R record $G/rollout_handarm_long.trxb problem=dexsyn kind=handarm hidden=8 frames=24 tiles=1 programs=0 keep_dr=0
# LeastSquaresSolver iterates of the dexsyn_small problem (values only): reference engine work, restated Eigen CG
$R record $G/sq_small.trxb problem=dexsyn frames=3 hidden=4 tiles=2 extra_fixed=4 sq_steps=3 programs=0 keep_dr=0
# the dexsyn_small problem with collision and friction-cone penalties (dexlearn.h:66-131,219-233,324-362), values only:
# reference operators, gradients and engine; the Bullet call inside them answered through the hook of
# oracle/ref_build/shim/tractor_collision/collision_decl.h by csrc/collision/trx_collide.h (parity of that query unpinned)
$R record $G/dexsyn_collision.trxb problem=dexsyn frames=3 hidden=4 tiles=2 extra_fixed=4 collision=1 friction=1 gd_steps=10 programs=0 gd_keep=last
# ... and without the floor / object pair (the object rests on the floor face to face: non-unique witnesses), for exact
# parity of the device, whose rounding differs from the host's
$R record $G/dexsyn_collision_nofloor.trxb problem=dexsyn frames=3 hidden=4 tiles=2 extra_fixed=4 collision=1 friction=1 floor_pair=0 gd_steps=10 programs=0 gd_keep=last
# constraint operators in every mode (single instructions), and the project / barrier programs buildConstraints derives
# for a small bounded least-squares tape (core/constraints.h, gradients.cpp:764-997)
$R constraint_tests $G/constraint_ops_8d.trxb suffix=8d
$R constraint_tests $G/constraint_ops_d.trxb suffix=d
$R constrained $G/constrained_small.trxb
ls -la $G
