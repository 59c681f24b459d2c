/*
 * csm_oracle.c -- stage-2 oracle: dependency-free restatement of the vendored
 * Cartographer fast correlative scan matcher in registration/2d (plus the
 * author's Grid2D additions).  TEST INFRASTRUCTURE ONLY (see gloc_oracle.h).
 *
 * PARITY PINNED against the reference's own code: registration/2d/*.cpp,
 * 3d/probability_values.cpp and 3d/point_cloud.cpp compile UNMODIFIED into
 * oracle/_ref/libcsm_ref.so (oracle/csm_ref.cpp; oracle/shim/ stands in for the
 * Eigen / glog / OpenCV / boost / ceres headers this image lacks), and
 * tests/test_oracle_csm_ref.py holds this file equal to it: precomputation grids,
 * value codec, SearchParameters, rotated + discretised scans, and the result of
 * MatchWithSearchParameters / MatchFullSubmap (score, candidate, pose) bit for bit;
 * tests/golden/csm_*.npz are minted from the reference.  What remains a
 * restatement is Eigen's own arithmetic (not under /root/reference, version
 * unpinned by the reference's CMake): Quaternion(AngleAxis),
 * QuaternionBase::_transformVector and Transform * vector as in Eigen 3.3/3.4,
 * stated once in oracle/shim/Eigen/eigen_shim.h (general form) and once here
 * (specialised to rotations about z); the two agree bit for bit.  Every function
 * cites the source lines it restates.
 *
 * Compile with -ffp-contract=off and no -ffast-math: float32 arithmetic must
 * stay un-fused and un-reassociated.
 */
#include "gloc_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------ value codec */

/* 3d/probability_values.h:64-67 */
static const float kMinProbability = 0.1f;
#define K_MAX_PROBABILITY (1.f - kMinProbability)
#define K_MIN_COST (1.f - K_MAX_PROBABILITY)
#define K_MAX_COST (1.f - kMinProbability)

float gloc_oracle_min_cost(void) { return K_MIN_COST; }
float gloc_oracle_max_cost(void) { return K_MAX_COST; }

/* SlowValueToBoundedFloat, 3d/probability_values.cpp:27-36; the table built
 * at :38-52 repeats it for values with bit 15 (update marker) set. */
float gloc_oracle_value_to_cost(uint16_t value) {
  const uint16_t v = (uint16_t)(value & 32767u);
  const float lower_bound = K_MIN_COST, upper_bound = K_MAX_COST;
  if (v == 0) return K_MAX_COST; /* kUnknownCorrespondenceValue -> kMax (:59-63) */
  const float kScale = (upper_bound - lower_bound) / 32766.f;
  return (float)v * kScale + (lower_bound - kScale);
}

/* port.h:41-43 RoundToInt = std::lround */
static int round_to_int_f(float x) { return (int)lroundf(x); }
static int round_to_int_d(double x) { return (int)lround(x); }

/* BoundedFloatToValue, 3d/probability_values.h:32-44 */
uint16_t gloc_oracle_cost_to_value(float cost) {
  const float lo = K_MIN_COST, hi = K_MAX_COST;
  float c = cost;
  if (c > hi) c = hi; /* common::Clamp, 3d/math.h:31-39 */
  if (c < lo) c = lo;
  const int value = round_to_int_f((c - lo) * (32766.f / (hi - lo))) + 1;
  return (uint16_t)value;
}

/* ComputeCellValue, 2d/fast_correlative_scan_matcher_2d.cpp:184-190 with
 * min_score_/max_score_ from :118-119 */
uint8_t gloc_oracle_cell_value(float probability) {
  const float min_score = 1.f - K_MAX_COST;
  const float max_score = 1.f - K_MIN_COST;
  const int cell_value =
      round_to_int_f((probability - min_score) * (255.f / (max_score - min_score)));
  return (uint8_t)cell_value;
}

/* Grid2D::GetCorrespondenceCost, 2d/grid_2d.cpp:86-90 (+ ToFlatIndex :168-171) */
static float grid_cost(const uint16_t* cells, int nx, int ny, int x, int y) {
  if (x < 0 || y < 0 || x >= nx || y >= ny) return K_MAX_COST;
  return gloc_oracle_value_to_cost(cells[(size_t)nx * (size_t)y + (size_t)x]);
}

static float grid_prob(const uint16_t* cells, int nx, int ny, int x, int y) {
  /* "1.f - std::abs(grid.GetCorrespondenceCost(...))", fast_..._2d.cpp:130-131 */
  return 1.f - fabsf(grid_cost(cells, nx, ny, x, y));
}

void gloc_oracle_level1_from_cells(const uint16_t* cells, int nx, int ny,
                                   uint8_t* out) {
  for (int y = 0; y < ny; ++y)
    for (int x = 0; x < nx; ++x)
      out[(size_t)nx * y + x] =
          gloc_oracle_cell_value(grid_prob(cells, nx, ny, x, y));
}

/* SlidingWindowMaximum, 2d/fast_correlative_scan_matcher_2d.cpp:43-76, on a
 * ring buffer instead of std::deque. */
typedef struct {
  float* buf;
  int cap, head, size;
} swm;

static void swm_init(swm* s, int cap) {
  s->buf = (float*)malloc(sizeof(float) * (size_t)cap);
  s->cap = cap;
  s->head = 0;
  s->size = 0;
}
static void swm_reset(swm* s) { s->head = 0; s->size = 0; }
static float swm_back(const swm* s) {
  return s->buf[(s->head + s->size - 1) % s->cap];
}
static void swm_add(swm* s, float v) { /* :45-51 */
  while (s->size > 0 && v > swm_back(s)) s->size--;
  s->buf[(s->head + s->size) % s->cap] = v;
  s->size++;
}
static void swm_remove(swm* s, float v) { /* :53-61 */
  if (v == s->buf[s->head]) {
    s->head = (s->head + 1) % s->cap;
    s->size--;
  }
}
static float swm_max(const swm* s) { return s->buf[s->head]; } /* :63-68 */

/* PrecomputationGrid2D ctor, 2d/fast_correlative_scan_matcher_2d.cpp:112-182 */
void gloc_oracle_precomp_from_cells(const uint16_t* cells, int nx, int ny,
                                    int width, uint8_t* out) {
  const int wide_nx = nx + width - 1;
  const int wide_ny = ny + width - 1;
  const int stride = wide_nx;
  float* intermediate = (float*)malloc(sizeof(float) * (size_t)wide_nx * (size_t)ny);
  swm cur;
  swm_init(&cur, width + 2);
  (void)wide_ny;
  for (int y = 0; y != ny; ++y) { /* :127-151 */
    swm_reset(&cur);
    swm_add(&cur, grid_prob(cells, nx, ny, 0, y));
    for (int x = -width + 1; x != 0; ++x) {
      intermediate[x + width - 1 + y * stride] = swm_max(&cur);
      if (x + width < nx) swm_add(&cur, grid_prob(cells, nx, ny, x + width, y));
    }
    for (int x = 0; x < nx - width; ++x) {
      intermediate[x + width - 1 + y * stride] = swm_max(&cur);
      swm_remove(&cur, grid_prob(cells, nx, ny, x, y));
      swm_add(&cur, grid_prob(cells, nx, ny, x + width, y));
    }
    for (int x = (nx - width > 0 ? nx - width : 0); x != nx; ++x) {
      intermediate[x + width - 1 + y * stride] = swm_max(&cur);
      swm_remove(&cur, grid_prob(cells, nx, ny, x, y));
    }
  }
  for (int x = 0; x != wide_nx; ++x) { /* :155-181 */
    swm_reset(&cur);
    swm_add(&cur, intermediate[x]);
    for (int y = -width + 1; y != 0; ++y) {
      out[x + (y + width - 1) * stride] = gloc_oracle_cell_value(swm_max(&cur));
      if (y + width < ny) swm_add(&cur, intermediate[x + (y + width) * stride]);
    }
    for (int y = 0; y < ny - width; ++y) {
      out[x + (y + width - 1) * stride] = gloc_oracle_cell_value(swm_max(&cur));
      swm_remove(&cur, intermediate[x + y * stride]);
      swm_add(&cur, intermediate[x + (y + width) * stride]);
    }
    for (int y = (ny - width > 0 ? ny - width : 0); y != ny; ++y) {
      out[x + (y + width - 1) * stride] = gloc_oracle_cell_value(swm_max(&cur));
      swm_remove(&cur, intermediate[x + y * stride]);
    }
  }
  free(cur.buf);
  free(intermediate);
}

/* Statement implemented by the GPU path: cell (x0,y0), x0 in [-w+1,nx),
 * y0 in [-w+1,ny), = max of level1 over [x0,x0+w) x [y0,y0+w) clipped to the
 * grid (header comment fast_..._2d.h:58-60 + ctor semantics). */
void gloc_oracle_precomp_from_level1(const uint8_t* level1, int nx, int ny,
                                     int width, uint8_t* out) {
  const int wide_nx = nx + width - 1;
  const int wide_ny = ny + width - 1;
  uint8_t* rowmax = (uint8_t*)malloc((size_t)wide_nx * (size_t)ny);
  for (int y = 0; y < ny; ++y)
    for (int lx = 0; lx < wide_nx; ++lx) {
      const int x0 = lx - width + 1;
      uint8_t m = 0;
      for (int x = (x0 < 0 ? 0 : x0); x < x0 + width && x < nx; ++x) {
        const uint8_t v = level1[(size_t)nx * y + x];
        if (v > m) m = v;
      }
      rowmax[(size_t)wide_nx * y + lx] = m;
    }
  for (int ly = 0; ly < wide_ny; ++ly) {
    const int y0 = ly - width + 1;
    for (int lx = 0; lx < wide_nx; ++lx) {
      uint8_t m = 0;
      for (int y = (y0 < 0 ? 0 : y0); y < y0 + width && y < ny; ++y) {
        const uint8_t v = rowmax[(size_t)wide_nx * y + lx];
        if (v > m) m = v;
      }
      out[(size_t)wide_nx * ly + lx] = m;
    }
  }
  free(rowmax);
}

/* ------------------------------------------------------- search parameters */

/* SearchParameters(double,double,PointCloud,double), correlative_scan_matcher_2d.cpp:27-55 */
void gloc_oracle_search_params(double linear_window, double angular_window,
                               const float* pts, int n_pts, double resolution,
                               int* n_lin, int* n_ang, double* ang_step) {
  float max_scan_range = (float)(3.f * resolution); /* :34 */
  for (int i = 0; i < n_pts; ++i) {
    const float x = pts[3 * i], y = pts[3 * i + 1];
    const float range = sqrtf(x * x + y * y); /* point.head<2>().norm() */
    if (range > max_scan_range) max_scan_range = range; /* std::max(range, max) */
  }
  const double kSafetyMargin = 1. - 1e-3;
  const float r2 = max_scan_range * max_scan_range; /* common::Pow2<float> */
  *ang_step =
      kSafetyMargin * acos(1. - (resolution * resolution) / (2. * r2)); /* :40-42 */
  *n_ang = (int)ceil(angular_window / *ang_step); /* :43-44 */
  *n_lin = (int)ceil(linear_window / resolution); /* :47-48 */
}

/* GridToVirtualPointCloud, fast_..._2d.cpp:78-95 */
int gloc_oracle_grid_to_points(const uint16_t* cells, int nx, int ny,
                               double resolution, double ox, double oy,
                               float* pts, int capacity) {
  int n = 0;
  for (int i = 0; i < nx; ++i) {
    for (int j = 0; j < ny; ++j) {
      if (grid_cost(cells, nx, ny, i, j) < 0.11) { /* float promoted to double */
        if (pts != NULL && n < capacity) {
          pts[3 * n] = (float)(ox + i * resolution);
          pts[3 * n + 1] = (float)(oy + j * resolution);
          pts[3 * n + 2] = 0.f;
        }
        ++n;
      }
    }
  }
  return n;
}

/* ----------------------------------------------------- rotate + discretise */

/* Eigen: Quaternionf(AngleAxisf(theta, UnitZ)) then QuaternionBase::_transformVector
 *   ha = 0.5f*theta; w = cos(ha); vec = sin(ha)*axis
 *   uv = vec.cross(v); uv += uv; return v + w*uv + vec.cross(uv);
 * used by Rigid3f::Rotation (3d/rigid_transform.h:134-136) and operator*
 * (:209-214, "+ translation" with translation == 0) via TransformPointCloud
 * (3d/point_cloud.cpp:23-31). */
static void rotate_z(float theta, const float* in, int n, float* out) {
  const float ha = 0.5f * theta;
  const float qw = cosf(ha);
  const float s = sinf(ha);
  const float qx = s * 0.f, qy = s * 0.f, qz = s * 1.f;
  for (int i = 0; i < n; ++i) {
    const float vx = in[3 * i], vy = in[3 * i + 1], vz = in[3 * i + 2];
    float ux = qy * vz - qz * vy;
    float uy = qz * vx - qx * vz;
    float uz = qx * vy - qy * vx;
    ux += ux;
    uy += uy;
    uz += uz;
    const float cx = qy * uz - qz * uy;
    const float cy = qz * ux - qx * uz;
    const float cz = qx * uy - qy * ux;
    out[3 * i] = ((vx + qw * ux) + cx) + 0.f;
    out[3 * i + 1] = ((vy + qw * uy) + cy) + 0.f;
    out[3 * i + 2] = ((vz + qw * uz) + cz) + 0.f;
  }
}

void gloc_oracle_discretize(const float* pts, int n_pts, double init_x,
                            double init_y, double init_yaw, int n_ang,
                            double ang_step, double resolution, double max_x,
                            double max_y, int32_t* out_cells) {
  const int S = 2 * n_ang + 1;
  float* p0 = (float*)malloc(sizeof(float) * 3 * (size_t)(n_pts > 0 ? n_pts : 1));
  float* ps = (float*)malloc(sizeof(float) * 3 * (size_t)(n_pts > 0 ? n_pts : 1));
  /* fast_..._2d.cpp:278-283: rotate by initial_rotation.cast<float>().angle() */
  rotate_z((float)init_yaw, pts, n_pts, p0);
  /* Eigen::Translation2f(double, double): converted to float at the call (:287-288) */
  const float tx = (float)init_x, ty = (float)init_y;
  /* correlative_scan_matcher_2d.cpp:99-107: delta_theta accumulated in double */
  double delta_theta = -n_ang * ang_step;
  for (int s = 0; s < S; ++s, delta_theta += ang_step) {
    rotate_z((float)delta_theta, p0, n_pts, ps);
    for (int i = 0; i < n_pts; ++i) {
      /* Affine2f(translation) * point.head<2>()  (:119-121) */
      const float wx = ps[3 * i] + tx;
      const float wy = ps[3 * i + 1] + ty;
      /* MapLimits::GetCellIndex, map_limits.h:69-76 (double arithmetic) */
      const int cx = round_to_int_d((max_y - (double)wy) / resolution - 0.5);
      const int cy = round_to_int_d((max_x - (double)wx) / resolution - 0.5);
      out_cells[2 * ((size_t)s * n_pts + i)] = cx;
      out_cells[2 * ((size_t)s * n_pts + i) + 1] = cy;
    }
  }
  free(p0);
  free(ps);
}

/* ------------------------------------------------------------ the matcher */

typedef struct {
  int w, wide_nx, wide_ny;
  uint8_t* cells;
} level_grid;

typedef struct {
  int scan, xo, yo;
  float score;
  int order; /* generation order, for the stable sort */
} cand;

typedef struct {
  int min_x, max_x, min_y, max_y;
} bounds;

typedef struct {
  const level_grid* levels;
  int depth;
  const int32_t* cells; /* S x P x 2 */
  int n_pts, S;
  const bounds* lb;
  long long n_scored;
} match_ctx;

/* PrecomputationGrid2D::GetValue, fast_..._2d.h:68-83 */
static inline int level_value(const level_grid* g, int x, int y) {
  const int lx = x + g->w - 1, ly = y + g->w - 1;
  if ((unsigned)lx >= (unsigned)g->wide_nx || (unsigned)ly >= (unsigned)g->wide_ny)
    return 0;
  return g->cells[lx + ly * g->wide_nx];
}

/* ToScore, fast_..._2d.h:86-88 */
static inline float to_score(float value) {
  const float min_score = 1.f - K_MAX_COST;
  const float max_score = 1.f - K_MIN_COST;
  return min_score + value * ((max_score - min_score) / 255.f);
}

static int cand_desc(const void* a, const void* b) {
  const cand* ca = (const cand*)a;
  const cand* cb = (const cand*)b;
  if (ca->score > cb->score) return -1;
  if (ca->score < cb->score) return 1;
  return (ca->order > cb->order) - (ca->order < cb->order);
}

/* ScoreCandidates, fast_..._2d.cpp:372-391 (std::sort -> stable order) */
static void score_candidates(match_ctx* c, const level_grid* g, cand* cs, int n) {
  for (int i = 0; i < n; ++i) {
    int sum = 0;
    const int32_t* pts = c->cells + 2 * (size_t)cs[i].scan * c->n_pts;
    for (int p = 0; p < c->n_pts; ++p)
      sum += level_value(g, pts[2 * p] + cs[i].xo, pts[2 * p + 1] + cs[i].yo);
    cs[i].score = to_score(sum / (float)c->n_pts);
    cs[i].order = i;
  }
  c->n_scored += n;
  qsort(cs, (size_t)n, sizeof(cand), cand_desc);
}

/* BranchAndBound, fast_..._2d.cpp:393-438 */
static cand branch_and_bound(match_ctx* c, const cand* cs, int n, int cand_depth,
                             float min_score) {
  if (cand_depth == 0) return cs[0]; /* :399-402 */
  cand best = {0, 0, 0, min_score, 0}; /* :406-407 */
  for (int i = 0; i < n; ++i) {
    if (cs[i].score <= min_score) break; /* :409-411 */
    cand hi[4];
    int nh = 0;
    const int half_width = 1 << (cand_depth - 1);
    const bounds* b = &c->lb[cs[i].scan];
    for (int xi = 0; xi < 2; ++xi) { /* :414-429 */
      const int xoff = xi * half_width;
      if (cs[i].xo + xoff > b->max_x) break;
      for (int yi = 0; yi < 2; ++yi) {
        const int yoff = yi * half_width;
        if (cs[i].yo + yoff > b->max_y) break;
        hi[nh].scan = cs[i].scan;
        hi[nh].xo = cs[i].xo + xoff;
        hi[nh].yo = cs[i].yo + yoff;
        ++nh;
      }
    }
    score_candidates(c, &c->levels[cand_depth - 1], hi, nh); /* :430-432 */
    const cand sub = branch_and_bound(c, hi, nh, cand_depth - 1, best.score);
    if (best.score < sub.score) best = sub; /* std::max(best, sub), :433-436 */
  }
  return best;
}

static void build_levels(const uint8_t* level1, int nx, int ny, int depth,
                         level_grid* levels) {
  /* PrecomputationGridStack2D, fast_..._2d.cpp:192-207: widths 1,2,4,... */
  for (int i = 0; i < depth; ++i) {
    const int w = 1 << i;
    levels[i].w = w;
    levels[i].wide_nx = nx + w - 1;
    levels[i].wide_ny = ny + w - 1;
    levels[i].cells =
        (uint8_t*)malloc((size_t)levels[i].wide_nx * (size_t)levels[i].wide_ny);
    gloc_oracle_precomp_from_level1(level1, nx, ny, w, levels[i].cells);
  }
}

int gloc_oracle_csm_match(const uint8_t* level1, int nx, int ny,
                          double resolution, double max_x, double max_y,
                          int depth, const float* pts, int n_pts,
                          double init_x, double init_y, double init_yaw,
                          int n_lin, int n_ang, double ang_step,
                          float min_score, int mode,
                          gloc_oracle_match_result* out) {
  memset(out, 0, sizeof(*out));
  out->score = min_score;
  if (depth < 1 || n_pts < 1 || nx < 1 || ny < 1) return -1;
  const int S = 2 * n_ang + 1;
  int32_t* cells = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)S * (size_t)n_pts);
  gloc_oracle_discretize(pts, n_pts, init_x, init_y, init_yaw, n_ang, ang_step,
                         resolution, max_x, max_y, cells);
  /* SearchParameters test ctor (:57-71) + ShrinkToFit (:73-91) */
  bounds* lb = (bounds*)malloc(sizeof(bounds) * (size_t)S);
  for (int s = 0; s < S; ++s) {
    int minbx = 0, minby = 0, maxbx = 0, maxby = 0;
    const int32_t* sp = cells + 2 * (size_t)s * n_pts;
    for (int p = 0; p < n_pts; ++p) {
      const int x = sp[2 * p], y = sp[2 * p + 1];
      if (-x < minbx) minbx = -x;
      if (-y < minby) minby = -y;
      if (nx - 1 - x > maxbx) maxbx = nx - 1 - x;
      if (ny - 1 - y > maxby) maxby = ny - 1 - y;
    }
    lb[s].min_x = -n_lin > minbx ? -n_lin : minbx;
    lb[s].max_x = n_lin < maxbx ? n_lin : maxbx;
    lb[s].min_y = -n_lin > minby ? -n_lin : minby;
    lb[s].max_y = n_lin < maxby ? n_lin : maxby;
  }
  level_grid* levels = (level_grid*)calloc((size_t)depth, sizeof(level_grid));
  match_ctx ctx = {levels, depth, cells, n_pts, S, lb, 0};
  cand best = {0, 0, 0, min_score, 0};
  if (mode == 0) {
    build_levels(level1, nx, ny, depth, levels);
    /* GenerateLowestResolutionCandidates, fast_..._2d.cpp:334-370 */
    const int step = 1 << (depth - 1);
    size_t num = 0;
    for (int s = 0; s < S; ++s) {
      const int cx = (lb[s].max_x - lb[s].min_x + step) / step;
      const int cy = (lb[s].max_y - lb[s].min_y + step) / step;
      num += (size_t)(cx > 0 ? cx : 0) * (size_t)(cy > 0 ? cy : 0);
    }
    cand* cs = (cand*)malloc(sizeof(cand) * (num > 0 ? num : 1));
    int n = 0;
    for (int s = 0; s < S; ++s)
      for (int xo = lb[s].min_x; xo <= lb[s].max_x; xo += step)
        for (int yo = lb[s].min_y; yo <= lb[s].max_y; yo += step) {
          cs[n].scan = s;
          cs[n].xo = xo;
          cs[n].yo = yo;
          ++n;
        }
    if (n > 0) {
      /* ComputeLowestResolutionCandidates (:322-332) then BranchAndBound (:304-306) */
      score_candidates(&ctx, &levels[depth - 1], cs, n);
      best = branch_and_bound(&ctx, cs, n, depth - 1, min_score);
    }
    free(cs);
  } else {
    build_levels(level1, nx, ny, 1, levels);
    for (int s = 0; s < S; ++s) {
      const int32_t* sp = cells + 2 * (size_t)s * n_pts;
      for (int xo = lb[s].min_x; xo <= lb[s].max_x; ++xo)
        for (int yo = lb[s].min_y; yo <= lb[s].max_y; ++yo) {
          int sum = 0;
          for (int p = 0; p < n_pts; ++p)
            sum += level_value(&levels[0], sp[2 * p] + xo, sp[2 * p + 1] + yo);
          const float sc = to_score(sum / (float)n_pts);
          ++ctx.n_scored;
          if (sc > best.score) { /* first maximum in (scan, x, y) order */
            best.scan = s;
            best.xo = xo;
            best.yo = yo;
            best.score = sc;
          }
        }
    }
  }
  out->n_scored = ctx.n_scored;
  if (best.score > min_score) { /* fast_..._2d.cpp:311-319 */
    out->found = 1;
    out->score = best.score;
    out->scan_index = best.scan;
    out->x_offset = best.xo;
    out->y_offset = best.yo;
    /* Candidate2D, correlative_scan_matcher_2d.h:81-86 */
    const double cx = -best.yo * resolution;
    const double cy = -best.xo * resolution;
    const double orientation = (best.scan - n_ang) * ang_step;
    out->pose_x = init_x + cx;
    out->pose_y = init_y + cy;
    out->pose_yaw = init_yaw + orientation; /* Rotation2Dd product = angle sum */
  }
  for (int i = 0; i < depth; ++i) free(levels[i].cells);
  free(levels);
  free(lb);
  free(cells);
  return out->found;
}

/* MatchFullSubmap, fast_..._2d.cpp:249-268 */
int gloc_oracle_csm_match_full_submap(const uint8_t* level1, int nx, int ny,
                                      double resolution, double max_x,
                                      double max_y, int depth, const float* pts,
                                      int n_pts, float min_score, int mode,
                                      gloc_oracle_match_result* out) {
  int n_lin, n_ang;
  double step;
  gloc_oracle_search_params(25 * resolution, M_PI, pts, n_pts, resolution,
                            &n_lin, &n_ang, &step);
  /* center = max - 0.5*res*(num_x_cells, num_y_cells)  (:258-261) */
  const double cx = max_x - 0.5 * resolution * nx;
  const double cy = max_y - 0.5 * resolution * ny;
  return gloc_oracle_csm_match(level1, nx, ny, resolution, max_x, max_y, depth,
                               pts, n_pts, cx, cy, 0.0, n_lin, n_ang, step,
                               min_score, mode, out);
}

typedef struct {
  const uint8_t* const* grids;
  int nx, ny;
  double resolution, max_x, max_y;
  int depth;
  const float* const* pts;
  const int* n_pts;
  const double* init;
  int n_pairs, n_lin, n_ang;
  double ang_step;
  float min_score;
  int mode;
  int t, nthreads;
  gloc_oracle_match_result* out;
} csm_job;

static void* csm_worker(void* p) {
  csm_job* j = (csm_job*)p;
  for (int i = j->t; i < j->n_pairs; i += j->nthreads)
    gloc_oracle_csm_match(j->grids[i], j->nx, j->ny, j->resolution, j->max_x,
                          j->max_y, j->depth, j->pts[i], j->n_pts[i],
                          j->init[3 * i], j->init[3 * i + 1], j->init[3 * i + 2],
                          j->n_lin, j->n_ang, j->ang_step, j->min_score,
                          j->mode, &j->out[i]);
  return NULL;
}

void gloc_oracle_csm_match_batch_mt(const uint8_t* const* grids, int nx, int ny,
                                    double resolution, double max_x,
                                    double max_y, int depth,
                                    const float* const* pts, const int* n_pts,
                                    const double* init_xyyaw, int n_pairs,
                                    int n_lin, int n_ang, double ang_step,
                                    float min_score, int mode, int nthreads,
                                    gloc_oracle_match_result* out) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n_pairs) nthreads = n_pairs;
  if (n_pairs <= 0) return;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  csm_job* jobs = (csm_job*)malloc(sizeof(csm_job) * (size_t)nthreads);
  for (int t = 0; t < nthreads; ++t) {
    jobs[t] = (csm_job){grids, nx, ny, resolution, max_x, max_y, depth, pts,
                        n_pts, init_xyyaw, n_pairs, n_lin, n_ang, ang_step,
                        min_score, mode, t, nthreads, out};
    pthread_create(&th[t], NULL, csm_worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
}
