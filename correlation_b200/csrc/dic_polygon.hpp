// dic_polygon.hpp -- host-side geometry of blob domains (tiny: O(vertices^2 + rows)).
//
// Blob membership follows the reference CPU engine (polygon_class.cpp), not its CUDA functor:
//   triangulate()      :224-281  O'Rourke ear clipping on a CCW-oriented simple loop
//   trianglePoints()   :283-337  split at y_mid into two flat triangles
//   flatTrianglePoints :339-403  rows j in [ceil(ya), ceil(yb)), cols i in [ceil(x_small(j)), ceil(x_big(j)))
// The pixels themselves are never touched here: each flat triangle yields row spans
// (y, x_begin, x_end) in the reference's emission order, and the device expands them
// (expand_spans_kernel). Shared edges are evaluated by each triangle with its own fp32 line,
// exactly like the reference, so its occasional duplicate / missing edge pixel is reproduced.
#pragma once
#include <cmath>
#include <vector>

namespace dic {

struct HostSpan { int y, xb, xe; };

class BlobPolygon {
  struct V { float x, y; bool ear; int next, prev; };
  std::vector<V> v_;
  int head_ = 0, count_ = 0;
  bool bad_ = false;
  std::vector<int> tri_; // 3 vertex ids per triangle, emission order

  float area2(int a, int b, int c) const { // :49-57
    return (v_[b].x - v_[a].x) * (v_[c].y - v_[a].y) - (v_[c].x - v_[a].x) * (v_[b].y - v_[a].y);
  }
  bool left(int a, int b, int c) const { return area2(a, b, c) > 0.f; }
  bool leftOn(int a, int b, int c) const { return area2(a, b, c) >= 0.f; }
  bool collinear(int a, int b, int c) const { return area2(a, b, c) == 0.f; }
  bool intersectProp(int a, int b, int c, int d) const { // :109-118
    if (collinear(a, b, c) || collinear(a, b, d) || collinear(b, d, a) || collinear(c, d, b)) return false;
    return (!left(a, b, c) ^ !left(a, b, d)) && (!left(c, d, a) ^ !left(c, d, b));
  }
  bool between(int a, int b, int c) const { // :120-139
    if (!collinear(a, b, c)) return false;
    if (v_[a].x != v_[b].x)
      return (v_[a].x <= v_[c].x && v_[c].x <= v_[b].x) || (v_[a].x >= v_[c].x && v_[c].x >= v_[b].x);
    return (v_[a].y <= v_[c].y && v_[c].y <= v_[b].y) || (v_[a].y >= v_[c].y && v_[c].y >= v_[b].y);
  }
  bool intersect(int a, int b, int c, int d) const { // :141-152
    return intersectProp(a, b, c, d) || between(a, b, c) || between(a, b, d) || between(c, d, a) ||
           between(c, d, b);
  }
  bool diagonalIE(int a, int b) const { // :154-173
    int c = head_;
    do {
      int c1 = v_[c].next;
      if (c != a && c1 != a && c != b && c1 != b && intersect(a, b, c, c1)) return false;
      c = c1;
    } while (c != head_);
    return true;
  }
  bool inCone(int a, int b) const { // :175-187
    int a1 = v_[a].next, a0 = v_[a].prev;
    if (leftOn(a, a1, a0)) return left(a, b, a0) && left(b, a, a1);
    return !(leftOn(a, b, a1) && leftOn(b, a, a0));
  }
  bool diagonal(int a, int b) const { return inCone(a, b) && inCone(b, a) && diagonalIE(a, b); }
  float areaPolyTwice() const { // :68-81
    float sum = 0.f;
    int a = v_[head_].next;
    do {
      sum += area2(head_, a, v_[a].next);
      a = v_[a].next;
    } while (v_[a].next != head_);
    return sum;
  }
  bool simpleLoop() const { // :195-222
    if (count_ < 4) return true;
    int ol = head_;
    do {
      int orr = v_[ol].next;
      int il = v_[orr].next;
      do {
        int ir = v_[il].next;
        if (intersect(ol, orr, il, ir)) return false;
        il = ir;
      } while (il != head_ && il != v_[ol].prev);
      ol = orr;
    } while (ol != v_[v_[head_].prev].prev);
    return true;
  }

  static bool line(float x1, float y1, float x2, float y2, float &dxdy, float &x0) { // :405-416
    float den = y2 - y1;
    if (den == 0) return true;
    dxdy = (x2 - x1) / den;
    x0 = x1 - dxdy * y1;
    return false;
  }
  static void flat(std::vector<HostSpan> &out, float x1, float y1, float x2, float y2, float x3,
                   float y3) { // :339-403
    int dy = (int)(std::floor((double)y3) - std::floor((double)y1));
    int dx = (int)(std::floor((double)x2) - std::floor((double)x1));
    if (dx == 0 || dy == 0) return;
    float xs = dx > 0 ? x1 : x2, ys = dx > 0 ? y1 : y2;
    float xb = dx > 0 ? x2 : x1, yb = dx > 0 ? y2 : y1;
    float ds = 0.f, db = 0.f, x0s = 0.f, x0b = 0.f;
    line(xs, ys, x3, y3, ds, x0s);
    line(xb, yb, x3, y3, db, x0b);
    int j0 = dy > 0 ? (int)std::ceil((double)y1) : (int)std::ceil((double)y3);
    int j1 = dy > 0 ? (int)std::ceil((double)y3) : (int)std::ceil((double)y1);
    for (int j = j0; j < j1; ++j) {
      volatile float es = ds * (float)j; // keep mul and add as two fp32 roundings
      volatile float eb = db * (float)j;
      int i0 = (int)std::ceil((double)(float)(es + x0s));
      int i1 = (int)std::ceil((double)(float)(eb + x0b));
      if (i1 > i0) out.push_back(HostSpan{j, i0, i1});
    }
  }
  void triangleSpans(std::vector<HostSpan> &out, int a, int b, int c) const { // :283-337
    const V *ymax, *ymid, *ymin;
    const V &A = v_[a], &B = v_[b], &C = v_[c];
    if (A.y > B.y) {
      if (B.y > C.y) { ymax = &A; ymid = &B; ymin = &C; }
      else if (C.y > A.y) { ymax = &C; ymid = &A; ymin = &B; }
      else { ymax = &A; ymid = &C; ymin = &B; }
    } else {
      if (A.y > C.y) { ymax = &B; ymid = &A; ymin = &C; }
      else if (C.y > B.y) { ymax = &C; ymid = &B; ymin = &A; }
      else { ymax = &B; ymid = &C; ymin = &A; }
    }
    float dxdy, x0;
    if (line(ymin->x, ymin->y, ymax->x, ymax->y, dxdy, x0)) return;
    float newY = ymid->y;
    volatile float m = dxdy * newY;
    float newX = m + x0;
    flat(out, ymid->x, ymid->y, newX, newY, ymax->x, ymax->y);
    flat(out, ymid->x, ymid->y, newX, newY, ymin->x, ymin->y);
  }

public:
  BlobPolygon(const float *contour_xy, int n) {
    v_.resize(n > 0 ? n : 0);
    count_ = n;
    for (int i = 0; i < n; ++i)
      v_[i] = V{contour_xy[2 * i], contour_xy[2 * i + 1], false, (i + 1) % n, (i + n - 1) % n};
    if (n < 3 || !simpleLoop()) { bad_ = true; return; }
    if (areaPolyTwice() < 0) // reOrientPoly :83-97
      for (auto &q : v_) std::swap(q.next, q.prev);
    int w = head_;
    do { // earInit :37-47
      v_[w].ear = diagonal(v_[w].prev, v_[w].next);
      w = v_[w].next;
    } while (w != head_);
    while (count_ > 3) {
      int v2 = head_;
      bool found = false;
      do {
        if (v_[v2].ear) {
          int v3 = v_[v2].next, v4 = v_[v3].next, v1 = v_[v2].prev, v0 = v_[v1].prev;
          tri_.insert(tri_.end(), {v1, v2, v3});
          v_[v1].ear = diagonal(v0, v3);
          v_[v3].ear = diagonal(v1, v4);
          v_[v1].next = v3;
          v_[v3].prev = v1;
          head_ = v3;
          --count_;
          found = true;
          break;
        }
        v2 = v_[v2].next;
      } while (v2 != head_);
      if (!found) { bad_ = true; return; } // the reference would spin forever here
    }
    tri_.insert(tri_.end(), {v_[head_].prev, head_, v_[head_].next});
  }
  bool bad() const { return bad_; }
  int triangles() const { return (int)tri_.size() / 3; }
  std::vector<HostSpan> spans() const { // getInsidePoints :418-429, as spans
    std::vector<HostSpan> out;
    for (size_t t = 0; t + 2 < tri_.size(); t += 3) triangleSpans(out, tri_[t], tri_[t + 1], tri_[t + 2]);
    return out;
  }
};

} // namespace dic
