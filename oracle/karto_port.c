/* ORACLE / TEST INFRASTRUCTURE -- see karto_port.h.  Plain-C restatement of the reference
 * correlative scan matcher.  Each function cites the reference lines it follows
 * (M.cpp = /root/reference/lib/karto_sdk/src/Mapper.cpp, K.h / M.h / Math.h = the headers in
 * /root/reference/lib/karto_sdk/include/karto_sdk/).  Floating-point expressions keep the
 * reference's operand order; build with -ffp-contract=off (oracle/Makefile).
 */
#include "karto_port.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define KP_PI 3.14159265358979323846      /* Math.h:32 */
#define KP_2PI 6.28318530717958647692     /* Math.h:33 */
#define KP_PI_180 0.01745329251994329577  /* Math.h:35 */
#define KP_TOLERANCE 1e-06                /* Math.h:41 */
#define KP_INVALID_SCAN INT32_MAX         /* Math.h:47 */
#define KP_MAX_VARIANCE 500.0             /* M.cpp:52 */
#define KP_DISTANCE_PENALTY_GAIN 0.2      /* M.cpp:53 */
#define KP_ANGLE_PENALTY_GAIN 0.2         /* M.cpp:54 */
#define KP_OCCUPIED 100                   /* K.h:4380 GridStates_Occupied */

/* ---- Math.h ---------------------------------------------------------------------------- */
static double kp_round(double v) { return v >= 0.0 ? floor(v + 0.5) : ceil(v - 0.5); } /* Math.h:87 */
static int kp_double_equal(double a, double b)                                          /* Math.h:136 */
{
  double delta = a - b;
  return delta < 0.0 ? delta >= -KP_TOLERANCE : delta <= KP_TOLERANCE;
}
static double kp_square(double v) { return v * v; }
static double kp_max(double a, double b) { return a > b ? a : b; }  /* Math.h:112 Maximum */
static int kp_is_up_to(int32_t v, int32_t m) { return v >= 0 && v < m; } /* Math.h:149 */
static double kp_normalize_angle(double angle)                           /* Math.h:182-205 */
{
  while (angle < -KP_PI) {
    if (angle < -KP_2PI) angle += (uint32_t)(angle / -KP_2PI) * KP_2PI; else angle += KP_2PI;
  }
  while (angle > KP_PI) {
    if (angle > KP_2PI) angle -= (uint32_t)(angle / KP_2PI) * KP_2PI; else angle -= KP_2PI;
  }
  return angle;
}
static double kp_normalize_angle_difference(double minuend, double subtrahend) /* Math.h:215-226 */
{
  while (minuend - subtrahend < -KP_PI) minuend += KP_2PI;
  while (minuend - subtrahend > KP_PI) minuend -= KP_2PI;
  return minuend;
}
/* static_cast<kt_int32s>(double) as x86-64 cvttsd2si does it: NaN/Inf/out-of-range -> INT_MIN */
static int32_t kp_to_int(double r)
{
  if (!(r > -2147483649.0 && r < 2147483648.0)) return INT32_MIN;
  return (int32_t)r;
}

/* ---- grids ----------------------------------------------------------------------------- */
typedef struct {
  int32_t width, height, stride;          /* K.h:4636-4664 */
  int32_t roi_x, roi_y, roi_w, roi_h;     /* M.h:1204 */
  double scale, off_x, off_y;             /* CoordinateConverter K.h:4393-4560 */
  uint8_t * data;
  int32_t ksize;
  uint8_t * kernel;
  double smear;
} kp_cgrid;

typedef struct {
  int32_t width, height, stride;
  double scale, off_x, off_y;
  double * data;
} kp_dgrid;

struct kp_matcher {
  kp_params p;
  kp_cgrid g;
  kp_dgrid probs;
  /* lookup table (GridIndexLookup) */
  int32_t lut_angles, lut_n, lut_cap;
  int32_t * lut;
};

static double kp_resolution(double scale) { return 1.0 / scale; } /* K.h:4518 */

/* CoordinateConverter::WorldToGrid K.h:4421-4436 (flipY=false) */
static void kp_world_to_grid(double scale, double offx, double offy, double wx, double wy,
                             int32_t * gx, int32_t * gy)
{
  double gridX = (wx - offx) * scale;
  double gridY = (wy - offy) * scale;
  *gx = kp_to_int(kp_round(gridX));
  *gy = kp_to_int(kp_round(gridY));
}

static int32_t kp_half_kernel(double smear, double resolution) /* M.h:1275-1280 */
{
  return (int32_t)kp_round(2.0 * smear / resolution);
}

/* CorrelationGrid::CalculateKernel M.h:1213-1266 */
static int kp_calc_kernel(kp_cgrid * g)
{
  double resolution = kp_resolution(g->scale);
  const double min_dev = 0.5 * resolution, max_dev = 10 * resolution;
  if (!(g->smear >= min_dev && g->smear <= max_dev)) return -1; /* reference throws */
  g->ksize = 2 * kp_half_kernel(g->smear, resolution) + 1;
  g->kernel = (uint8_t *)malloc((size_t)g->ksize * g->ksize);
  int32_t half = g->ksize / 2;
  for (int32_t i = -half; i <= half; i++) {
    for (int32_t j = -half; j <= half; j++) {
      double d = hypot(i * resolution, j * resolution);
      double z = exp(-0.5 * pow(d / g->smear, 2));
      uint32_t kv = (uint32_t)kp_round(z * KP_OCCUPIED);
      g->kernel[(i + half) + g->ksize * (j + half)] = (uint8_t)kv;
    }
  }
  return 0;
}

void kp_point_readings(const double * ranges, int32_t n, const double sensor_pose[3],
                       double minimum_angle, double angular_resolution, double * points_xy)
{
  /* K.h:5660-5683: both branches compute the same point; unfiltered keeps every beam */
  for (int32_t i = 0; i < n; i++) {
    double r = ranges[i];
    double angle = sensor_pose[2] + minimum_angle + (uint32_t)i * angular_resolution;
    points_xy[2 * i] = sensor_pose[0] + (r * cos(angle));
    points_xy[2 * i + 1] = sensor_pose[1] + (r * sin(angle));
  }
}

kp_matcher * kp_create(const kp_params * p) /* M.cpp:477-522 */
{
  if (p->resolution <= 0) return NULL;
  if (p->search_size <= 0) return NULL;
  if (p->smear_deviation < 0) return NULL;
  if (p->range_threshold <= 0) return NULL;
  uint32_t side = (uint32_t)(kp_round(p->search_size / p->resolution) + 1);
  uint32_t margin = (uint32_t)ceil(p->range_threshold / p->resolution);
  int32_t grid_size = (int32_t)(side + 2 * margin);

  kp_matcher * m = (kp_matcher *)calloc(1, sizeof(kp_matcher));
  m->p = *p;
  /* CorrelationGrid::CreateGrid M.h:1099-1114, ctor M.h:1194-1208 */
  uint32_t border = (uint32_t)kp_half_kernel(p->smear_deviation, p->resolution) + 1;
  kp_cgrid * g = &m->g;
  g->width = (int32_t)(grid_size + border * 2);
  g->height = g->width;
  g->stride = (int32_t)(((size_t)g->width + 7) & ~(size_t)7); /* Math.h:234 AlignValue */
  g->scale = 1.0 / p->resolution;
  g->roi_x = (int32_t)border; g->roi_y = (int32_t)border; g->roi_w = grid_size; g->roi_h = grid_size;
  g->smear = p->smear_deviation;
  g->data = (uint8_t *)calloc((size_t)g->stride * g->height, 1);
  if (kp_calc_kernel(g) != 0) { free(g->data); free(m); return NULL; }
  /* search space probabilities Grid<double> M.cpp:513 */
  kp_dgrid * d = &m->probs;
  d->width = (int32_t)side; d->height = (int32_t)side;
  d->stride = (int32_t)(((size_t)d->width + 7) & ~(size_t)7);
  d->scale = 1.0 / p->resolution;
  d->data = (double *)calloc((size_t)d->stride * d->height, sizeof(double));
  return m;
}

void kp_destroy(kp_matcher * m)
{
  if (!m) return;
  free(m->g.data); free(m->g.kernel); free(m->probs.data); free(m->lut); free(m);
}

/* CorrelationGrid::GridIndex M.h:1122-1128 (ROI-relative -> linear) */
static int32_t kp_grid_index(const kp_cgrid * g, int32_t gx, int32_t gy)
{
  return (gx + g->roi_x) + (gy + g->roi_y) * g->stride;
}

int32_t kp_find_valid_points(const kp_scan * scan, const double viewpoint[2], double * out) /* M.cpp:1113-1164 */
{
  const double * pts = scan->points_xy;
  const double minSquareDistance = kp_square(0.1);
  int32_t trailing = 0, nout = 0;
  double fx = 0.0, fy = 0.0;
  int firstTime = 1;
  for (int32_t i = 0; i < scan->n; i++) {
    double cx = pts[2 * i], cy = pts[2 * i + 1];
    if (firstTime && !isnan(cx) && !isnan(cy)) { fx = cx; fy = cy; firstTime = 0; }
    double dx = fx - cx, dy = fy - cy;
    if (dx * dx + dy * dy > minSquareDistance) {
      double a = viewpoint[1] - fy;
      double b = fx - viewpoint[0];
      double c = fy * viewpoint[0] - fx * viewpoint[1];
      double ss = cx * a + cy * b + c;
      fx = cx; fy = cy;
      if (ss < 0.0) {
        trailing = i;
      } else {
        for (; trailing != i; ++trailing) { out[2 * nout] = pts[2 * trailing]; out[2 * nout + 1] = pts[2 * trailing + 1]; nout++; }
      }
    }
  }
  return nout;
}

/* CorrelationGrid::SmearPoint M.h:1152-1183 */
static void kp_smear_point(kp_cgrid * g, int32_t gx, int32_t gy)
{
  int32_t idx = kp_grid_index(g, gx, gy);
  if (g->data[idx] != KP_OCCUPIED) return;
  int32_t half = g->ksize / 2;
  for (int32_t j = -half; j <= half; j++) {
    uint8_t * adr = g->data + kp_grid_index(g, gx, gy + j);
    int32_t kc = half + g->ksize * (j + half);
    for (int32_t i = -half; i <= half; i++) {
      uint8_t kv = g->kernel[i + kc];
      if (kv > adr[i]) adr[i] = kv;
    }
  }
}

/* ScanMatcher::AddScan M.cpp:1073-1105 */
static void kp_add_scan(kp_matcher * m, const kp_scan * s, const double viewpoint[2], double * tmp)
{
  kp_cgrid * g = &m->g;
  int32_t nv = kp_find_valid_points(s, viewpoint, tmp);
  for (int32_t k = 0; k < nv; k++) {
    int32_t gx, gy;
    kp_world_to_grid(g->scale, g->off_x, g->off_y, tmp[2 * k], tmp[2 * k + 1], &gx, &gy);
    if (!kp_is_up_to(gx, g->roi_w) || !kp_is_up_to(gy, g->roi_h)) continue;
    int32_t idx = kp_grid_index(g, gx, gy);
    if (g->data[idx] == KP_OCCUPIED) continue;
    g->data[idx] = KP_OCCUPIED;
    kp_smear_point(g, gx, gy);
  }
}

void kp_raster(kp_matcher * m, const kp_scan * query, const kp_scan * base, int32_t nbase)
{
  kp_cgrid * g = &m->g;
  /* M.cpp:560-569 */
  g->off_x = query->sensor_pose[0] - (0.5 * (g->roi_w - 1) * kp_resolution(g->scale));
  g->off_y = query->sensor_pose[1] - (0.5 * (g->roi_h - 1) * kp_resolution(g->scale));
  /* AddScans M.cpp:1032-1045 */
  memset(g->data, 0, (size_t)g->stride * g->height);
  int32_t maxn = 0;
  for (int32_t i = 0; i < nbase; i++) if (base[i].n > maxn) maxn = base[i].n;
  double * tmp = (double *)malloc(sizeof(double) * 2 * (size_t)(maxn > 0 ? maxn : 1));
  double vp[2] = {query->sensor_pose[0], query->sensor_pose[1]};
  for (int32_t i = 0; i < nbase; i++) kp_add_scan(m, &base[i], vp, tmp);
  free(tmp);
}

/* GridIndexLookup::ComputeOffsets K.h:6797-6894 into m->lut */
static int32_t kp_compute_offsets(kp_matcher * m, const kp_scan * q, double angleCenter,
                                  double angleOffset, double angleResolution)
{
  const kp_cgrid * g = &m->g;
  uint32_t nAngles = (uint32_t)(kp_round(angleOffset * 2.0 / angleResolution) + 1);
  int32_t n = q->n;
  if ((int64_t)nAngles * n > m->lut_cap) {
    free(m->lut);
    m->lut_cap = (int32_t)(nAngles * (uint32_t)n);
    m->lut = (int32_t *)malloc(sizeof(int32_t) * (size_t)m->lut_cap);
  }
  m->lut_angles = (int32_t)nAngles; m->lut_n = n;
  /* Transform(sensorPose) K.h:2953, SetTransform K.h:3002-3024, Matrix3::FromAxisAngle K.h:2482-2511 */
  double m00, m01, m02, m10, m11, m12, tx, ty, th;
  const double * sp = q->sensor_pose;
  if (sp[0] == 0.0 && sp[1] == 0.0 && sp[2] == 0.0) {
    m00 = 1; m01 = 0; m02 = 0; m10 = 0; m11 = 1; m12 = 0; tx = 0; ty = 0; th = 0;
  } else {
    double radians = 0.0 - sp[2];
    double c = cos(radians), s = sin(radians), omc = 1.0 - c;
    double x = 0, y = 0, z = 1;
    double xyM = x * y * omc, xzM = x * z * omc, yzM = y * z * omc;
    double xS = x * s, yS = y * s, zS = z * s;
    m00 = x * x * omc + c; m01 = xyM - zS; m02 = xzM + yS;
    m10 = xyM + zS; m11 = y * y * omc + c; m12 = yzM - xS;
    tx = sp[0]; ty = sp[1]; th = sp[2] - 0.0;
  }
  double * lx = (double *)malloc(sizeof(double) * 2 * (size_t)(n > 0 ? n : 1));
  for (int32_t i = 0; i < n; i++) { /* InverseTransformPose K.h:2987-2994 on Pose2(point, 0.0) */
    double dx = q->points_xy[2 * i] - tx, dy = q->points_xy[2 * i + 1] - ty, dh = 0.0 - th;
    lx[2 * i] = m00 * dx + m01 * dy + m02 * dh;
    lx[2 * i + 1] = m10 * dx + m11 * dy + m12 * dh;
  }
  double startAngle = angleCenter - angleOffset;
  for (uint32_t a = 0; a < nAngles; a++) {
    double angle = startAngle + a * angleResolution;
    double cosine = cos(angle), sine = sin(angle);
    int32_t * out = m->lut + (size_t)a * n;
    for (int32_t i = 0; i < n; i++) {
      if (isnan(q->ranges[i]) || isinf(q->ranges[i])) { out[i] = KP_INVALID_SCAN; continue; }
      double ox = cosine * lx[2 * i] - sine * lx[2 * i + 1];
      double oy = sine * lx[2 * i] + cosine * lx[2 * i + 1];
      int32_t gx, gy;
      kp_world_to_grid(g->scale, g->off_x, g->off_y, ox + g->off_x, oy + g->off_y, &gx, &gy);
      out[i] = gx + (gy * g->stride); /* Grid<T>::GridIndex(.., false) K.h:4692 */
    }
  }
  free(lx);
  return (int32_t)nAngles;
}

int32_t kp_offsets(kp_matcher * m, const kp_scan * q, double c, double o, double r, int32_t * out)
{
  int32_t na = kp_compute_offsets(m, q, c, o, r);
  memcpy(out, m->lut, sizeof(int32_t) * (size_t)na * q->n);
  return na;
}

/* ScanMatcher::GetResponse M.cpp:1172-1208, integer numerator only */
static int64_t kp_response_sum(const kp_matcher * m, int32_t angleIndex, int32_t gridPositionIndex)
{
  const kp_cgrid * g = &m->g;
  const int32_t data_size = g->stride * g->height;
  const uint8_t * pByte = g->data + gridPositionIndex;
  const int32_t * off = m->lut + (size_t)angleIndex * m->lut_n;
  int64_t sum = 0;
  for (int32_t i = 0; i < m->lut_n; i++) {
    if (off[i] == KP_INVALID_SCAN) continue;
    int64_t pgi = (int64_t)gridPositionIndex + off[i];
    if (!(pgi >= 0 && pgi < data_size)) continue;
    sum += pByte[off[i]];
  }
  return sum;
}
static double kp_get_response(const kp_matcher * m, int32_t a, int32_t gpi, int64_t * sum_out)
{
  if (m->lut_n == 0) { if (sum_out) *sum_out = 0; return 0.0; }
  int64_t s = kp_response_sum(m, a, gpi);
  if (sum_out) *sum_out = s;
  double response = (double)s;
  response /= (double)((uint32_t)m->lut_n * (uint32_t)KP_OCCUPIED);
  return response;
}

/* ScanMatcher::ComputePositionalCovariance M.cpp:874-966 */
static void kp_positional_cov(kp_matcher * m, const double best[3], double bestResponse,
                              const double center[3], const double off[2], const double res[2],
                              double searchAngleResolution, double cov[9])
{
  memset(cov, 0, 9 * sizeof(double)); cov[0] = cov[4] = cov[8] = 1.0;
  if (bestResponse < KP_TOLERANCE) {
    cov[0] = KP_MAX_VARIANCE; cov[4] = KP_MAX_VARIANCE; cov[8] = 4 * kp_square(searchAngleResolution);
    return;
  }
  double aXX = 0, aXY = 0, aYY = 0, norm = 0;
  double dx = best[0] - center[0], dy = best[1] - center[1];
  double offsetX = off[0], offsetY = off[1];
  uint32_t nX = (uint32_t)(kp_round(offsetX * 2.0 / res[0]) + 1);
  double startX = -offsetX;
  uint32_t nY = (uint32_t)(kp_round(offsetY * 2.0 / res[1]) + 1);
  double startY = -offsetY;
  const kp_dgrid * d = &m->probs;
  for (uint32_t yI = 0; yI < nY; yI++) {
    double y = startY + yI * res[1];
    for (uint32_t xI = 0; xI < nX; xI++) {
      double x = startX + xI * res[0];
      int32_t gx, gy;
      kp_world_to_grid(d->scale, d->off_x, d->off_y, center[0] + x, center[1] + y, &gx, &gy);
      double response = d->data[gx + gy * d->stride];
      if (response >= (bestResponse - 0.1)) {
        norm += response;
        aXX += (kp_square(x - dx) * response);
        aXY += ((x - dx) * (y - dy) * response);
        aYY += (kp_square(y - dy) * response);
      }
    }
  }
  if (norm > KP_TOLERANCE) {
    double vXX = aXX / norm, vXY = aXY / norm, vYY = aYY / norm;
    double vTT = 4 * kp_square(searchAngleResolution);
    double minXX = 0.1 * kp_square(res[0]), minYY = 0.1 * kp_square(res[1]);
    vXX = kp_max(vXX, minXX); vYY = kp_max(vYY, minYY);
    double mult = 1.0 / bestResponse;
    cov[0] = vXX * mult; cov[1] = vXY * mult; cov[3] = vXY * mult; cov[4] = vYY * mult; cov[8] = vTT;
  }
  if (kp_double_equal(cov[0], 0.0)) cov[0] = KP_MAX_VARIANCE;
  if (kp_double_equal(cov[4], 0.0)) cov[4] = KP_MAX_VARIANCE;
}

/* ScanMatcher::ComputeAngularCovariance M.cpp:977-1025 */
static void kp_angular_cov(kp_matcher * m, const double best[3], double bestResponse,
                           const double center[3], double searchAngleOffset,
                           double searchAngleResolution, double cov[9])
{
  const kp_cgrid * g = &m->g;
  double bestAngle = kp_normalize_angle_difference(best[2], center[2]);
  int32_t gx, gy;
  kp_world_to_grid(g->scale, g->off_x, g->off_y, best[0], best[1], &gx, &gy);
  int32_t gridIndex = kp_grid_index(g, gx, gy);
  uint32_t nAngles = (uint32_t)(kp_round(searchAngleOffset * 2 / searchAngleResolution) + 1);
  double startAngle = center[2] - searchAngleOffset;
  double norm = 0.0, acc = 0.0;
  for (uint32_t a = 0; a < nAngles; a++) {
    double angle = startAngle + a * searchAngleResolution;
    double response = kp_get_response(m, (int32_t)a, gridIndex, NULL);
    if (response >= (bestResponse - 0.1)) {
      norm += response;
      acc += (kp_square(angle - bestAngle) * response);
    }
  }
  if (norm > KP_TOLERANCE) {
    if (acc < KP_TOLERANCE) acc = kp_square(searchAngleResolution);
    acc /= norm;
  } else {
    acc = 1000 * kp_square(searchAngleResolution);
  }
  cov[8] = acc;
}

double kp_correlate(kp_matcher * m, const kp_scan * q, const double center[3], const double sp_off[2],
                    const double sp_res[2], double searchAngleOffset, double searchAngleResolution,
                    int32_t doPenalize, int32_t fine, double mean[3], double cov[9],
                    int32_t * sums, int32_t sums_cap, int32_t dims[3]) /* M.cpp:712-862 */
{
  kp_cgrid * g = &m->g;
  kp_dgrid * d = &m->probs;
  kp_compute_offsets(m, q, center[2], searchAngleOffset, searchAngleResolution);
  if (!fine) {
    memset(d->data, 0, sizeof(double) * (size_t)d->stride * d->height);
    d->off_x = center[0] - sp_off[0];
    d->off_y = center[1] - sp_off[1];
  }
  uint32_t nX = (uint32_t)(kp_round(sp_off[0] * 2.0 / sp_res[0]) + 1);
  double startX = -sp_off[0];
  uint32_t nY = (uint32_t)(kp_round(sp_off[1] * 2.0 / sp_res[1]) + 1);
  double startY = -sp_off[1];
  uint32_t nAngles = (uint32_t)(kp_round(searchAngleOffset * 2.0 / searchAngleResolution) + 1);
  uint32_t total = nX * nY * nAngles;
  if (dims) { dims[0] = (int32_t)nX; dims[1] = (int32_t)nY; dims[2] = (int32_t)nAngles; }
  double * resp = (double *)malloc(sizeof(double) * total);
  double * px = (double *)malloc(sizeof(double) * total);
  double * py = (double *)malloc(sizeof(double) * total);
  double * ph = (double *)malloc(sizeof(double) * total);
  /* operator()(y) M.cpp:641-694 for every row */
  for (uint32_t yI = 0; yI < nY; yI++) {
    double y = startY + yI * sp_res[1];
    double newPositionY = center[1] + y;
    double squareY = kp_square(y);
    for (uint32_t xI = 0; xI < nX; xI++) {
      double x = startX + xI * sp_res[0];
      double newPositionX = center[0] + x;
      double squareX = kp_square(x);
      int32_t gx, gy;
      kp_world_to_grid(g->scale, g->off_x, g->off_y, newPositionX, newPositionY, &gx, &gy);
      int32_t gridIndex = kp_grid_index(g, gx, gy);
      double startAngle = center[2] - searchAngleOffset;
      for (uint32_t a = 0; a < nAngles; a++) {
        double angle = startAngle + a * searchAngleResolution;
        int64_t s;
        double response = kp_get_response(m, (int32_t)a, gridIndex, &s);
        if (doPenalize && (kp_double_equal(response, 0.0) == 0)) {
          double squaredDistance = squareX + squareY;
          double distancePenalty = 1.0 - (KP_DISTANCE_PENALTY_GAIN * squaredDistance / m->p.distance_variance_penalty);
          distancePenalty = kp_max(distancePenalty, m->p.minimum_distance_penalty);
          double squaredAngleDistance = kp_square(angle - center[2]);
          double anglePenalty = 1.0 - (KP_ANGLE_PENALTY_GAIN * squaredAngleDistance / m->p.angle_variance_penalty);
          anglePenalty = kp_max(anglePenalty, m->p.minimum_angle_penalty);
          response *= (distancePenalty * anglePenalty);
        }
        uint32_t k = (yI * nX + xI) * nAngles + a;
        resp[k] = response; px[k] = newPositionX; py[k] = newPositionY; ph[k] = kp_normalize_angle(angle);
        if (sums && (int32_t)k < sums_cap) sums[k] = (int32_t)s;
      }
    }
  }
  /* M.cpp:775-800 */
  double bestResponse = -1;
  for (uint32_t i = 0; i < total; i++) {
    bestResponse = kp_max(bestResponse, resp[i]);
    if (!fine) {
      int32_t gx, gy;
      kp_world_to_grid(d->scale, d->off_x, d->off_y, px[i], py[i], &gx, &gy);
      /* reference throws on an out-of-range cell (M.cpp:786-796); cannot happen for valid params */
      if (kp_is_up_to(gx, d->width) && kp_is_up_to(gy, d->height)) {
        double * ptr = d->data + gx + gy * d->stride;
        *ptr = kp_max(resp[i], *ptr);
      }
    }
  }
  /* M.cpp:802-829 */
  double ax = 0.0, ay = 0.0, thetaX = 0.0, thetaY = 0.0;
  int32_t count = 0;
  for (uint32_t i = 0; i < total; i++) {
    if (kp_double_equal(resp[i], bestResponse)) {
      ax += px[i]; ay += py[i];
      thetaX += cos(ph[i]); thetaY += sin(ph[i]);
      count++;
    }
  }
  double avg[3] = {0, 0, 0};
  if (count > 0) {
    ax /= count; ay /= count; thetaX /= count; thetaY /= count;
    avg[0] = ax; avg[1] = ay; avg[2] = atan2(thetaY, thetaX);
  }
  free(resp); free(px); free(py); free(ph);
  if (!fine) kp_positional_cov(m, avg, bestResponse, center, sp_off, sp_res, searchAngleResolution, cov);
  else kp_angular_cov(m, avg, bestResponse, center, searchAngleOffset, searchAngleResolution, cov);
  mean[0] = avg[0]; mean[1] = avg[1]; mean[2] = avg[2];
  if (bestResponse > 1.0) bestResponse = 1.0;
  return bestResponse;
}

double kp_match(kp_matcher * m, const kp_scan * query, const kp_scan * base, int32_t nbase,
                int32_t doPenalize, int32_t doRefine, double mean[3], double cov[9]) /* M.cpp:534-639 */
{
  kp_cgrid * g = &m->g;
  memset(cov, 0, 9 * sizeof(double));
  if (query->n == 0) {
    mean[0] = query->sensor_pose[0]; mean[1] = query->sensor_pose[1]; mean[2] = query->sensor_pose[2];
    cov[0] = KP_MAX_VARIANCE; cov[4] = KP_MAX_VARIANCE;
    cov[8] = 4 * kp_square(m->p.coarse_angle_resolution);
    return 0.0;
  }
  kp_raster(m, query, base, nbase);
  double res = kp_resolution(g->scale);
  double dimX = m->probs.width, dimY = m->probs.height;
  double coarseOff[2] = {0.5 * (dimX - 1) * res, 0.5 * (dimY - 1) * res};
  double coarseRes[2] = {2 * res, 2 * res};
  double center[3] = {query->sensor_pose[0], query->sensor_pose[1], query->sensor_pose[2]};
  double best = kp_correlate(m, query, center, coarseOff, coarseRes, m->p.coarse_search_angle_offset,
                             m->p.coarse_angle_resolution, doPenalize, 0, mean, cov, NULL, 0, NULL);
  if (m->p.use_response_expansion) {
    if (kp_double_equal(best, 0.0)) {
      double newOff = m->p.coarse_search_angle_offset;
      for (uint32_t i = 0; i < 3; i++) {
        newOff += 20 * KP_PI_180;
        best = kp_correlate(m, query, center, coarseOff, coarseRes, newOff, m->p.coarse_angle_resolution,
                            doPenalize, 0, mean, cov, NULL, 0, NULL);
        if (kp_double_equal(best, 0.0) == 0) break;
      }
    }
  }
  if (doRefine) {
    double fineOff[2] = {coarseRes[0] * 0.5, coarseRes[1] * 0.5};
    double fineRes[2] = {res, res};
    double c2[3] = {mean[0], mean[1], mean[2]};
    best = kp_correlate(m, query, c2, fineOff, fineRes, 0.5 * m->p.coarse_angle_resolution,
                        m->p.fine_search_angle_offset, doPenalize, 1, mean, cov, NULL, 0, NULL);
  }
  return best;
}

const uint8_t * kp_grid(kp_matcher * m, int32_t info[9], double off[2])
{
  const kp_cgrid * g = &m->g;
  info[0] = g->width; info[1] = g->height; info[2] = g->stride; info[3] = g->roi_x; info[4] = g->roi_y;
  info[5] = g->roi_w; info[6] = g->roi_h; info[7] = g->stride * g->height; info[8] = g->ksize;
  off[0] = g->off_x; off[1] = g->off_y;
  return g->data;
}
const uint8_t * kp_kernel(kp_matcher * m) { return m->g.kernel; }

/* ======== occupancy grid: karto::OccupancyGrid::CreateFromScans (K.h:5946-5961) ============ */
struct kp_occupancy {
  int32_t width, height, stride;   /* K.h:4636-4640 */
  double scale, off_x, off_y;
  uint8_t * cells;                 /* Grid<kt_int8u>  */
  uint32_t * pass, * hits;         /* K.h:6292-6297   */
};

/* Grid<T>::TraceLine K.h:4874-4927: every visited valid cell gets ++ */
static void kp_trace_line(kp_occupancy * g, int32_t x0, int32_t y0, int32_t x1, int32_t y1)
{
  int steep = abs(y1 - y0) > abs(x1 - x0);
  int32_t t, deltaX, deltaY, error = 0, ystep, y, x;
  if (steep) { t = x0; x0 = y0; y0 = t; t = x1; x1 = y1; y1 = t; }
  if (x0 > x1) { t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; }
  deltaX = x1 - x0;
  deltaY = abs(y1 - y0);
  y = y0;
  ystep = y0 < y1 ? 1 : -1;
  for (x = x0; x <= x1; x++) {
    int32_t px = steep ? y : x, py = steep ? x : y;
    error += deltaY;
    if (2 * error >= deltaX) { y += ystep; error -= deltaX; }
    if (kp_is_up_to(px, g->width) && kp_is_up_to(py, g->height)) g->pass[px + py * g->stride]++;
  }
}

kp_occupancy * kp_occupancy_create(const kp_scan * scans, int32_t n, double resolution, double range_threshold,
                                   double minimum_range, double maximum_range, uint32_t min_pass_through,
                                   double occupancy_threshold)
{
  double minx = 999999999999999999.99999, miny = minx, maxx = -minx, maxy = -minx;   /* K.h:2845-2849 */
  kp_occupancy * g;
  int32_t s, i;
  size_t cells, c;
  if (n <= 0) return NULL;                                                            /* K.h:5950-5952 */
  /* ComputeDimensions K.h:6082-6107 over the scans' bounding boxes (LocalizedRangeScan::Update K.h:5692-5698) */
  for (s = 0; s < n; ++s) {
    const kp_scan * sc = &scans[s];
#define KP_BOX(x, y) do { if ((x) < minx) minx = (x); if ((y) < miny) miny = (y); if ((x) > maxx) maxx = (x); if ((y) > maxy) maxy = (y); } while (0)
    KP_BOX(sc->sensor_pose[0], sc->sensor_pose[1]);
    for (i = 0; i < sc->n; ++i) {
      double r = sc->ranges[i];
      if (r >= minimum_range && r <= range_threshold) KP_BOX(sc->points_xy[2 * i], sc->points_xy[2 * i + 1]);   /* InRange Math.h:123 */
    }
#undef KP_BOX
  }
  g = (kp_occupancy *)calloc(1, sizeof(*g));
  g->scale = 1.0 / resolution;
  g->width = kp_to_int(kp_round((maxx - minx) * g->scale));
  g->height = kp_to_int(kp_round((maxy - miny) * g->scale));
  g->stride = (g->width + 7) & ~7;                                                    /* AlignValue K.h:4640 */
  g->off_x = minx; g->off_y = miny;
  cells = (size_t)g->stride * (size_t)g->height;
  g->cells = (uint8_t *)calloc(cells ? cells : 1, 1);
  g->pass = (uint32_t *)calloc(cells ? cells : 1, sizeof(uint32_t));
  g->hits = (uint32_t *)calloc(cells ? cells : 1, sizeof(uint32_t));
  /* AddScan K.h:6139-6182 + RayTrace K.h:6193-6229 */
  for (s = 0; s < n; ++s) {
    const kp_scan * sc = &scans[s];
    double sx = sc->sensor_pose[0], sy = sc->sensor_pose[1];
    for (i = 0; i < sc->n; ++i) {
      double r = sc->ranges[i], px = sc->points_xy[2 * i], py = sc->points_xy[2 * i + 1];
      int valid_end = r < (range_threshold - KP_TOLERANCE);
      int32_t fx, fy, tx, ty;
      if (r <= minimum_range || r >= maximum_range || isnan(r)) continue;
      if (r >= range_threshold) {
        double ratio = range_threshold / r, dx = px - sx, dy = py - sy;
        px = sx + ratio * dx;
        py = sy + ratio * dy;
      }
      kp_world_to_grid(g->scale, g->off_x, g->off_y, sx, sy, &fx, &fy);
      kp_world_to_grid(g->scale, g->off_x, g->off_y, px, py, &tx, &ty);
      kp_trace_line(g, fx, fy, tx, ty);
      if (valid_end && kp_is_up_to(tx, g->width) && kp_is_up_to(ty, g->height)) {
        g->pass[tx + ty * g->stride]++;
        g->hits[tx + ty * g->stride]++;
      }
    }
  }
  /* Update K.h:6259-6274 / UpdateCell K.h:6241-6254 */
  for (c = 0; c < cells; ++c) {
    if (g->pass[c] > min_pass_through) {
      double ratio = (double)g->hits[c] / (double)g->pass[c];
      g->cells[c] = ratio > occupancy_threshold ? 100 : 255;   /* GridStates K.h:4379-4381 */
    }
  }
  return g;
}

void kp_occupancy_destroy(kp_occupancy * g)
{
  if (!g) return;
  free(g->cells); free(g->pass); free(g->hits); free(g);
}

void kp_occupancy_info(const kp_occupancy * g, int32_t info[3], double offset[2])
{
  info[0] = g->width; info[1] = g->height; info[2] = g->stride;
  offset[0] = g->off_x; offset[1] = g->off_y;
}
const uint8_t * kp_occupancy_cells(const kp_occupancy * g) { return g->cells; }
const uint32_t * kp_occupancy_pass(const kp_occupancy * g) { return g->pass; }
const uint32_t * kp_occupancy_hits(const kp_occupancy * g) { return g->hits; }
