"""Algorithmic fp64 work of a recorded objective: operations per timestep*rollout summed over the tapes.

SURVEY.md section 8(d) estimates F_alg from shapes; this module replaces the estimate by a sum over the
instruction stream of x_prog + x_prep + x_bprop: (instructions of each operator) x (fp64 operations of its
body).  The per-operator counts are read off the operator bodies (csrc/vm_ops.def, csrc/vm_math.cuh, which
restate tractor/include/tractor/core/ops.h, geometry/ops.h, dexopt/src/ops.h): add / sub / mul / compare-select
count 1, a multiply-add 2, division, square root and every libm call 1.
"""
from __future__ import annotations

import collections
import re
from typing import Dict

import numpy as np

from . import trxfile

QROT, QMUL, CROSS, DOT, QNORM = 30, 28, 9, 5, 14
QPLUS = 3 + QNORM + QMUL + QNORM
QAA = 6
PMUL = QROT + 3 + QMUL
PMULV = QROT + 3

# base operator name -> (compute, prepare, reverse) fp64 operations per lane
OPS = {
    "move": (0, 0, 0), "zero": (0, 0, 0), "add": (1, 0, 0), "sub": (1, 0, 1), "minus": (1, 0, 1),
    "mul": (1, 0, 2), "div": (1, 1, 5), "madd": (2, 0, 2), "tanh": (1, 2, 1), "exp": (1, 0, 1), "log": (1, 1, 1),
    "sin": (1, 1, 1), "cos": (1, 2, 1), "square": (1, 1, 1), "sqrt": (1, 2, 1), "relu": (1, 0, 0),
    "acos": (3, 6, 1), "sincos": (2, 1, 3),
    "add_fg_vec3": (3, 0, 0), "sub_fg_vec3": (3, 0, 3), "minus_fg_vec3": (3, 0, 3),
    "mul_fg_vec3_s": (3, 0, 8), "mul_fg_s_vec3": (3, 0, 8), "dot_fg_vec3": (DOT, 0, 6), "cross_fg_vec3": (CROSS, 0, 18),
    "fg_vec3_pack": (0, 0, 0), "fg_vec3_unpack": (0, 0, 0), "move_fg_vec3": (0, 0, 0), "move_fg_pose": (0, 0, 0),
    "move_fg_quat": (0, 0, 0), "move_fg_twist": (0, 0, 0), "move_fg_mat3": (0, 0, 0),
    "zero_fg_vec3": (0, 0, 0), "zero_fg_pose": (0, 0, 0), "zero_fg_quat": (0, 0, 0), "zero_fg_twist": (0, 0, 0),
    "zero_fg_mat3": (0, 0, 0), "add_fg_twist": (6, 0, 0), "sub_fg_twist": (6, 0, 6),
    "mul_fg_quat_vec3": (QROT, 0, QROT + CROSS + QROT + 3), "mul_fg_quat": (QMUL, 0, QROT + 3),
    "fg_quat_inverse": (3, 0, 3), "fg_quat_residual": (4, 0, 0), "fg_angle_axis_quat": (QAA, 0, DOT + 3),
    "mul_fg_pose": (PMUL, QROT + 3, CROSS + QROT + 3 + QROT), "mul_fg_pose_vec3": (PMULV, QROT + 3, CROSS + QROT),
    "fg_pose_angle_axis_pose": (QAA + QMUL, 3 + QROT, QROT + DOT + 3), "fg_angle_axis_pose": (QAA, 0, DOT + 3),
    "fg_pose_translation": (0, 0, 0), "fg_pose_orientation": (0, 0, 0), "fg_translation_pose": (0, 0, 0),
    "fg_orientation_pose": (0, 0, 0), "mul_fg_mat3_vec3": (18, 0, 27), "add_fg_quat_vec3": (QPLUS, 0, 0),
    "add_fg_pose_twist": (QPLUS + 6 + CROSS, 0, 0), "fg_pose_residual": (3 + QMUL + 3 + 4 + 3, 0, 6),
    "add_fg_mat3": (9, 0, 0), "batch": (0, 0, 7), "dot4add": (8, 0, 8), "gate": (0, 0, 1),
}
MODES = {"prepare": 1, "reverse": 2, "forward": None}


def split_name(name: str):
    """[mode_]base_<d|8d> -> (mode index or None, base)"""
    base = re.sub(r"_(8d|d)$", "", name)
    for mode, idx in MODES.items():
        if base.startswith(mode + "_"):
            return idx, base[len(mode) + 1:]
    return 0, base


def histogram(view: trxfile.ProgramView) -> Dict[str, int]:
    """instructions per operator name"""
    counts = collections.Counter()
    code = view.code
    n_args = [len(op.args) for op in view.ops]
    i = 0
    while i < len(code):
        oi = int(code[i])
        counts[view.ops[oi].name] += 1
        i += 1 + n_args[oi]
    return dict(counts)


def tape_flops(view: trxfile.ProgramView, dense_shapes=None):
    """(fp64 operations per lane of one execution of the tape, names of operators without a count)"""
    total, unknown = 0, set()
    n_dense = collections.Counter()
    for name, n in histogram(view).items():
        mode, base = split_name(name)
        lanes = max((a[2] for a in next(op for op in view.ops if op.name == name).args), default=1)
        if base in ("x_dense", "x_zero_block", "x_muld"):
            if base == "x_muld":
                total += 2 * n
            elif base == "x_dense":
                n_dense[mode] += n
            continue
        if lanes == 1:
            continue  # lane-independent instructions (weight decay rows): amortised over the rollouts
        if base not in OPS or mode is None:
            unknown.add(name)
            continue
        total += OPS[base][mode] * n
    if n_dense and dense_shapes:
        # the recorder emits the layers in order, frame after frame: instruction k uses shape k mod n_layers
        per_eval = sum(a * b for a, b in dense_shapes)
        bias = sum(b for _, b in dense_shapes)
        n_layers = len(dense_shapes)
        total += (n_dense[0] // n_layers) * (2 * per_eval + bias)   # y = xW + b
        total += (n_dense[2] // n_layers) * (4 * per_eval + bias)   # gx = gy W^T, gW += x^T gy, gb += gy
    return total, sorted(unknown)


def objective_flops_per_unit(build, options: dict, dense_shapes):
    """fp64 operations of x_prog + x_prep + x_bprop per timestep*rollout: difference of two horizons
    (so that the one-off start-up and terminal parts drop out)"""
    f1, f2 = 2, 4
    totals = []
    unknown = set()
    for frames in (f1, f2):
        bundle = build(frames=frames, fuse_dense=1, programs="prog,prep,bprop", **options)
        t = 0
        for name in ("prog", "prep", "bprop"):
            n, u = tape_flops(bundle.view(name), dense_shapes)
            t += n
            unknown.update(u)
        totals.append(t)
    return (totals[1] - totals[0]) / float(f2 - f1), sorted(unknown)
