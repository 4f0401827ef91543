// trx_shape_table.h -- host side of the collision shape table: builds the records csrc/collision/trx_collide.h reads.
// Shared by the library (trx_engine.cu: trx_collision_shape_create / _pair_create) and the test-only host emulation.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/trx_b200.h"
#include "trx_collide.h"

namespace trxcol {

struct ShapeTable {
  std::vector<double> heap{0.0};  // offset 0 is never a record: a zero handle is invalid
  std::vector<uint64_t> shapes, pairs;

  bool isShape(uint64_t h) const { return std::find(shapes.begin(), shapes.end(), h) != shapes.end(); }
  bool isPair(uint64_t h) const { return std::find(pairs.begin(), pairs.end(), h) != pairs.end(); }

  // returns an error text, empty on success (unsupported = kinds the reference also rejects, dexlearn.h:212-216)
  std::string addShape(const trx_shape_desc &d, uint64_t &handle, bool &unsupported) {
    unsupported = false;
    std::vector<double> rec(REC_DATA, 0.0);
    rec[REC_KIND] = d.kind;
    switch (d.kind) {
      case TRX_SHAPE_SPHERE:
        if (!(d.radius >= 0.0)) return "sphere radius";
        for (int i = 0; i < 3; i++) rec[REC_CENTER + i] = d.center[i];
        rec[REC_RADIUS] = d.radius;
        break;
      case TRX_SHAPE_CYLINDER:
        if (!(d.radius > 0.0) || !(d.length > 0.0)) return "cylinder radius / length";
        rec[REC_RADIUS] = d.radius;
        rec[REC_LENGTH] = d.length;
        break;
      case TRX_SHAPE_POLYHEDRON: {
        // shape.h:201-212: both lists non-empty, centre = mean of the points
        if (d.n_points == 0 || !d.points) return "points";
        if (d.n_planes == 0 || !d.planes) return "planes";
        double c[3] = {0, 0, 0};
        for (uint32_t i = 0; i < d.n_points; i++)
          for (int k = 0; k < 3; k++) c[k] += d.points[3 * i + k];
        for (int k = 0; k < 3; k++) rec[REC_CENTER + k] = c[k] * (1.0 / d.n_points);
        rec[REC_NPOINTS] = d.n_points;
        rec[REC_NPLANES] = d.n_planes;
        rec.insert(rec.end(), d.points, d.points + 3 * (size_t)d.n_points);
        rec.insert(rec.end(), d.planes, d.planes + 4 * (size_t)d.n_planes);
      } break;
      default:
        unsupported = true;
        return "collision shape not yet supported: kind " + std::to_string(d.kind);
    }
    for (double v : rec)
      if (!std::isfinite(v)) return "shape description is not finite";
    handle = heap.size();
    heap.insert(heap.end(), rec.begin(), rec.end());
    shapes.push_back(handle);
    return "";
  }

  std::string addPair(uint64_t a, uint64_t b, uint64_t &handle) {
    for (uint64_t h : {a, b})
      if (!isShape(h)) return "not a shape handle: " + std::to_string(h);
    handle = heap.size();
    heap.push_back((double)SHAPE_PAIR);
    heap.push_back((double)a);
    heap.push_back((double)b);
    pairs.push_back(handle);
    return "";
  }

  void clear() {
    heap.assign(1, 0.0);
    shapes.clear();
    pairs.clear();
  }
};

}  // namespace trxcol
