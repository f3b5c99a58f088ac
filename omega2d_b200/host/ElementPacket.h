// ElementPacket.h -- the ingest packet of the Omega2D containers, mirrored
// (reference src/ElementPacket.h:33-171): interleaved node coordinates x, node indices idx,
// per-element values val, element count nelem and element dimension ndim (0 points, 1 panels).
#pragma once
#ifndef OMEGA2D_REFERENCE_HEADERS
#include "O2DTypes.h"
#include <cassert>

template <class S>
class ElementPacket {
public:
  ElementPacket(std::vector<S> _x = std::vector<S>(), std::vector<Int> _idx = std::vector<Int>(),
                std::vector<S> _val = std::vector<S>(), size_t _nelem = 55, uint8_t _ndim = 55)
    : x(std::move(_x)), idx(std::move(_idx)), val(std::move(_val)), nelem(_nelem), ndim(_ndim) {}
  explicit ElementPacket(uint8_t _ndim) : ElementPacket(std::vector<S>(), std::vector<Int>(), std::vector<S>(), 0, _ndim) {}

  // append another packet of the same kind, re-basing its node indices (reference :78-129)
  void add(ElementPacket<S> _in) {
    assert(ndim == _in.ndim && "Should not be adding packets of different types");
    const Int nodecnt = (Int)(x.size() / Dimensions);
    for (auto& i : _in.idx) i += nodecnt;
    idx.insert(idx.end(), _in.idx.begin(), _in.idx.end());
    x.insert(x.end(), _in.x.begin(), _in.x.end());
    val.insert(val.end(), _in.val.begin(), _in.val.end());
    nelem += _in.nelem;
  }
  void translate(const S _x, const S _y) {
    for (size_t i = 0; i < x.size() / Dimensions; ++i) { x[i*Dimensions] += _x; x[i*Dimensions+1] += _y; }
  }

  std::vector<S> x;
  std::vector<Int> idx;
  std::vector<S> val;
  size_t nelem;
  uint8_t ndim;
};
#endif
