"""Reader for TRXB bundles (include/trx_file.h): serialised tractor tape programs + f64 arrays.

A bundle is the neutral form of tractor::Program objects
(reference: tractor/include/tractor/core/program.h:39-303) produced by the
reference-side dumper (oracle/ref_build/ref_tool.cpp) or by the product's own
tape builder.  Programs are kept as raw payload bytes (what
trx_program_create_from_blob takes) plus a light parsed view for tests.
"""
from __future__ import annotations

import lzma
import struct
from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np


@dataclass
class Port:
    address: int
    offset: int
    size: int
    lanes: int
    elem: int


@dataclass
class OpDesc:
    name: str
    args: List[tuple]  # (size, is_input, lanes, elem)


@dataclass
class ProgramView:
    lane_width: int
    memory_size: int
    ops: List[OpDesc]
    n_instructions: int
    code: np.ndarray
    inputs: List[Port]
    parameters: List[Port]
    outputs: List[Port]
    constants: List[Port]
    goals: List[tuple]
    const_data: bytes


@dataclass
class Bundle:
    programs: Dict[str, bytes] = field(default_factory=dict)
    arrays: Dict[str, np.ndarray] = field(default_factory=dict)
    texts: Dict[str, str] = field(default_factory=dict)

    def view(self, name: str) -> ProgramView:
        return parse_program(self.programs[name])

    def meta(self) -> Dict[str, str]:
        out = {}
        for line in self.texts.get("meta", "").splitlines():
            if "=" in line:
                k, v = line.split("=", 1)
                out[k] = v
        return out


class _Reader:
    def __init__(self, data: bytes):
        self.d = memoryview(data)
        self.p = 0

    def take(self, n: int) -> memoryview:
        if self.p + n > len(self.d):
            raise ValueError("trxfile: truncated")
        v = self.d[self.p:self.p + n]
        self.p += n
        return v

    def u8(self):
        return self.take(1)[0]

    def u16(self):
        return struct.unpack("<H", self.take(2))[0]

    def u32(self):
        return struct.unpack("<I", self.take(4))[0]

    def i32(self):
        return struct.unpack("<i", self.take(4))[0]

    def u64(self):
        return struct.unpack("<Q", self.take(8))[0]

    def s(self):
        return bytes(self.take(self.u16())).decode()


def _ports(r: _Reader) -> List[Port]:
    out = []
    for _ in range(r.u32()):
        a, o, sz, lanes, elem = struct.unpack("<QQIBB", r.take(22))
        out.append(Port(a, o, sz, lanes, elem))
    return out


def parse_program(payload: bytes) -> ProgramView:
    r = _Reader(payload)
    lane_width = r.u32()
    memory_size = r.u64()
    ops = []
    for _ in range(r.u32()):
        name = r.s()
        args = []
        for _ in range(r.u8()):
            size, is_in, lanes, elem = struct.unpack("<IBBB", r.take(7))
            args.append((size, is_in, lanes, elem))
        ops.append(OpDesc(name, args))
    n_inst = r.u64()
    n_code = r.u64()
    code = np.frombuffer(r.take(8 * n_code), dtype="<u8")
    inputs, parameters, outputs, constants = _ports(r), _ports(r), _ports(r), _ports(r)
    goals = []
    for _ in range(r.u32()):
        port = r.u64()
        prio = r.i32()
        goals.append((port, prio))
    const_data = bytes(r.take(r.u64()))
    return ProgramView(lane_width, memory_size, ops, n_inst, code, inputs, parameters, outputs, constants, goals,
                       const_data)


def encode_program(lane_width: int, memory_size: int, ops, instructions, inputs=(), parameters=(), outputs=(),
                   constants=(), const_data: bytes = b"") -> bytes:
    """program payload (the inverse of parse_program) from plain Python values:
    ops = [(name, [(size, is_input, lanes, elem), ...])], instructions = [(op index, [byte address, ...])],
    ports = [Port]; used to hand-write small tapes (tests of single operators)"""
    out = [struct.pack("<IQI", lane_width, memory_size, len(ops))]
    for name, args in ops:
        nb = name.encode()
        out.append(struct.pack("<H", len(nb)) + nb + struct.pack("<B", len(args)))
        for size, is_in, lanes, elem in args:
            out.append(struct.pack("<IBBB", size, is_in, lanes, elem))
    code = []
    for oi, addrs in instructions:
        code += [oi] + list(addrs)
    out.append(struct.pack("<QQ", len(instructions), len(code)))
    out.append(np.asarray(code, dtype="<u8").tobytes())
    for ports in (inputs, parameters, outputs, constants):
        out.append(struct.pack("<I", len(ports)))
        for q in ports:
            out.append(struct.pack("<QQIBB", q.address, q.offset, q.size, q.lanes, q.elem))
    out.append(struct.pack("<I", 0))  # goals
    out.append(struct.pack("<Q", len(const_data)) + const_data)
    return b"".join(out)


def loads(data: bytes) -> Bundle:
    r = _Reader(data)
    if bytes(r.take(8)) != b"TRXB0001":
        raise ValueError("trxfile: bad magic")
    b = Bundle()
    for _ in range(r.u32()):
        kind = r.u32()
        name = r.s()
        payload = r.take(r.u64())
        if kind == 1:
            b.programs[name] = bytes(payload)
        elif kind == 2:
            b.arrays[name] = np.frombuffer(payload, dtype="<f8").copy()
        elif kind == 3:
            b.texts[name] = bytes(payload).decode()
        else:
            raise ValueError("trxfile: unknown section kind %d" % kind)
    return b


def load(path: str) -> Bundle:
    with open(path, "rb") as f:
        data = f.read()
    if str(path).endswith(".xz"):
        data = lzma.decompress(data)
    return loads(data)


def wide_layout(ports: List[Port], lanes: int):
    """Element offsets of each port in the wide packed buffer (see include/trx_b200.h)."""
    out = []
    off = 0
    for p in ports:
        if p.lanes == 1:
            n = p.size // 8
            out.append((off, n, 1))
        else:
            comps = p.size // (8 * p.lanes)
            n = comps * lanes
            out.append((off, comps, lanes))
        off += n
    return out, off


def tiles_to_wide(ports: List[Port], tile_buffers: List[np.ndarray]) -> np.ndarray:
    """Reference-format packed buffers (one per 8-lane tile) -> wide layout for R = 8 * tiles lanes.

    Scalar-typed ports are taken from tile 0."""
    n_tiles = len(tile_buffers)
    lanes = 8 * n_tiles
    layout, total = wide_layout(ports, lanes)
    wide = np.zeros(total)
    for p, (off, comps, width) in zip(ports, layout):
        src0 = p.offset // 8
        if p.lanes == 1:
            wide[off:off + comps] = tile_buffers[0][src0:src0 + comps]
        else:
            for t, buf in enumerate(tile_buffers):
                block = buf[src0:src0 + comps * 8].reshape(comps, 8)
                wide[off:off + comps * lanes].reshape(comps, lanes)[:, t * 8:(t + 1) * 8] = block
    return wide


def wide_to_tiles(ports: List[Port], wide: np.ndarray, n_tiles: int) -> List[np.ndarray]:
    """Inverse of tiles_to_wide; scalar-typed ports are replicated into every tile buffer."""
    lanes = 8 * n_tiles
    layout, total = wide_layout(ports, lanes)
    assert wide.size == total
    tile_elems = sum(p.size for p in ports) // 8
    out = [np.zeros(tile_elems) for _ in range(n_tiles)]
    for p, (off, comps, width) in zip(ports, layout):
        dst0 = p.offset // 8
        if p.lanes == 1:
            for buf in out:
                buf[dst0:dst0 + comps] = wide[off:off + comps]
        else:
            block = wide[off:off + comps * lanes].reshape(comps, lanes)
            for t, buf in enumerate(out):
                buf[dst0:dst0 + comps * 8] = block[:, t * 8:(t + 1) * 8].reshape(-1)
    return out
