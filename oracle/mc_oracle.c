/* TEST INFRASTRUCTURE ONLY -- scalar CPU marching cubes used as the parity oracle
 * for the CUDA pipeline in sculptmate_b200/csrc/mcubes.cu.  Nothing under
 * sculptmate_b200/ links, loads or calls this file; only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY UNPINNED against the reference's real dependency: the reference calls
 * skimage.measure.marching_cubes (Lewiner) at
 * /root/reference/TripoSR/tsr/models/isosurface.py:46-48; scikit-image is a
 * third-party, version-unpinned dependency (requirements.txt:5) that is not
 * installed here, not in the wheelhouse and not vendored, and the reference has
 * no tests or golden meshes.  This file therefore *defines* the in-repo
 * algorithm (classic 256-case MC, table derived by tools/gen_mc_tables.py) and
 * follows the reference only for the wrapper semantics it can be pinned to:
 *   isosurface.py:45   level = -input.view(R,R,R); iso 0.0; array axes (x,y,z)
 *   isosurface.py:52   faces[:, [1,0,2]]  (flag SMB_MC_FLIP)
 *   isosurface.py:53   verts / (R-1)      (flag SMB_MC_DIV, IEEE fp32 division)
 *   system.py:185-189  scale_tensor(v,(0,1),(-r,r)) = v*fl(2r) + fl(-r)  (flag SMB_MC_AFFINE)
 *
 * Canonical output order (what "bit-exact" means for the CUDA path):
 *   vertices: for each x-plane i ascending: first the crossings on in-plane edges,
 *             visited by (j,k) row-major with the y-edge before the z-edge of a
 *             point, then the crossings on x-edges from plane i to i+1 by (j,k);
 *   triangles: by cell (i,j,k) row-major, then table order.
 *   A vertex on the edge (p -> p+e_axis) with values a=val(p), b=val(p+e_axis)
 *   sits at p + e_axis * (a / (a - b)), all in IEEE fp32.
 *   val(p) = (grid[p] - sub) * sign.
 * Slabs: with emit_last_plane = 0 the in-plane vertices of the last plane are
 *   numbered (after all of the slab's own vertices) but not stored or counted,
 *   so that concatenating slabs in order and adding the running vertex offset to
 *   each slab's faces reproduces the single-slab mesh bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "mc_tables_oracle.h"

#define SMB_MC_FLIP 1
#define SMB_MC_DIV 2
#define SMB_MC_AFFINE 4

typedef struct {
  int64_t nverts;      /* vertices stored by this slab */
  int64_t ntris;       /* triangles */
  int64_t npos;        /* grid points with val > 0 */
  int64_t nneg;        /* grid points with val < 0 */
  int64_t nverts_numbered; /* nverts + in-plane crossings of the last plane when not emitted */
} smb_oracle_mc_counts;

static inline float valf(const float *g, size_t idx, float sub, float sign) {
  volatile float d = g[idx] - sub; /* volatile: forbid contraction / reassociation */
  return d * sign;
}

static inline float xform(float v, int flags, float vdiv, float vmul, float vadd) {
  volatile float r = v;
  if (flags & SMB_MC_DIV) r = r / vdiv;
  if (flags & SMB_MC_AFFINE) {
    volatile float m = r * vmul;
    r = m + vadd;
  }
  return r;
}

/* grid: (nx,ny,nz) fp32, nz fastest.  verts (may be NULL): 3*nverts floats.
 * faces (may be NULL): 3*ntris int64.  Returns 0, or -1 on allocation failure. */
int smb_oracle_mc(const float *grid, int nx, int ny, int nz, float sub, float sign, int x_origin,
                  int emit_last_plane, int flags, float vdiv, float vmul, float vadd, float *verts,
                  int64_t *faces, smb_oracle_mc_counts *counts) {
  const size_t sy = (size_t)nz, sx = (size_t)ny * nz;
  const size_t npts = (size_t)nx * sx;
  /* vertex id per (point, axis); -1 = no crossing */
  int64_t *vid = (int64_t *)malloc(sizeof(int64_t) * 3 * npts);
  if (!vid) return -1;
  memset(vid, 0xff, sizeof(int64_t) * 3 * npts);
  int64_t nv = 0, nv_stored = 0, npos = 0, nneg = 0;

  for (int i = 0; i < nx; ++i) {
    const int store_plane = (i < nx - 1) || emit_last_plane;
    /* group 0: in-plane edges (y then z) */
    for (int j = 0; j < ny; ++j)
      for (int k = 0; k < nz; ++k) {
        size_t p = (size_t)i * sx + (size_t)j * sy + k;
        float a = valf(grid, p, sub, sign);
        if (a > 0.0f) ++npos;
        if (a < 0.0f) ++nneg;
        for (int axis = 1; axis <= 2; ++axis) {
          int ok = axis == 1 ? (j + 1 < ny) : (k + 1 < nz);
          if (!ok) continue;
          size_t q = p + (axis == 1 ? sy : 1);
          float b = valf(grid, q, sub, sign);
          if ((a > 0.0f) != (b > 0.0f)) {
            vid[3 * p + axis] = nv;
            if (store_plane) {
              if (verts) {
                volatile float den = a - b;
                volatile float t = a / den;
                float c[3] = {(float)(x_origin + i), (float)j, (float)k};
                volatile float moved = c[axis] + t;
                c[axis] = moved;
                for (int d = 0; d < 3; ++d) verts[3 * nv + d] = xform(c[d], flags, vdiv, vmul, vadd);
              }
              ++nv_stored;
            }
            ++nv;
          }
        }
      }
    /* group 1: x-edges from plane i to i+1 */
    if (i + 1 < nx)
      for (int j = 0; j < ny; ++j)
        for (int k = 0; k < nz; ++k) {
          size_t p = (size_t)i * sx + (size_t)j * sy + k;
          float a = valf(grid, p, sub, sign);
          float b = valf(grid, p + sx, sub, sign);
          if ((a > 0.0f) != (b > 0.0f)) {
            vid[3 * p + 0] = nv;
            if (verts) {
              volatile float den = a - b;
              volatile float t = a / den;
              volatile float moved = (float)(x_origin + i) + t;
              float c[3] = {moved, (float)j, (float)k};
              for (int d = 0; d < 3; ++d) verts[3 * nv + d] = xform(c[d], flags, vdiv, vmul, vadd);
            }
            ++nv_stored;
            ++nv;
          }
        }
  }

  int64_t nt = 0;
  for (int i = 0; i + 1 < nx; ++i)
    for (int j = 0; j + 1 < ny; ++j)
      for (int k = 0; k + 1 < nz; ++k) {
        size_t p = (size_t)i * sx + (size_t)j * sy + k;
        int cas = 0;
        for (int c = 0; c < 8; ++c) {
          size_t q = p + ((c >> 2) & 1) * sx + ((c >> 1) & 1) * sy + (c & 1);
          if (valf(grid, q, sub, sign) > 0.0f) cas |= 1 << c;
        }
        int n = SMB_MC_NTRI[cas];
        for (int t = 0; t < n; ++t) {
          if (faces) {
            int64_t id[3];
            for (int m = 0; m < 3; ++m) {
              int e = SMB_MC_TRI[cas][3 * t + m];
              size_t q = p + SMB_MC_EDGE_OWNER[e][0] * sx + SMB_MC_EDGE_OWNER[e][1] * sy +
                         SMB_MC_EDGE_OWNER[e][2];
              id[m] = vid[3 * q + SMB_MC_EDGE_OWNER[e][3]];
            }
            if (flags & SMB_MC_FLIP) {
              int64_t tmp = id[0];
              id[0] = id[1];
              id[1] = tmp;
            }
            faces[3 * nt + 0] = id[0];
            faces[3 * nt + 1] = id[1];
            faces[3 * nt + 2] = id[2];
          }
          ++nt;
        }
      }
  free(vid);
  if (counts) {
    counts->nverts = nv_stored;
    counts->ntris = nt;
    counts->npos = npos;
    counts->nneg = nneg;
    counts->nverts_numbered = nv;
  }
  return 0;
}

/* cube-case indices for every cell, (nx-1,ny-1,nz-1) uint8 -- used by the
 * "cube-case indices bit-exact" parity test. */
void smb_oracle_mc_cases(const float *grid, int nx, int ny, int nz, float sub, float sign,
                         unsigned char *cases) {
  const size_t sy = (size_t)nz, sx = (size_t)ny * nz;
  size_t o = 0;
  for (int i = 0; i + 1 < nx; ++i)
    for (int j = 0; j + 1 < ny; ++j)
      for (int k = 0; k + 1 < nz; ++k) {
        size_t p = (size_t)i * sx + (size_t)j * sy + k;
        int cas = 0;
        for (int c = 0; c < 8; ++c) {
          size_t q = p + ((c >> 2) & 1) * sx + ((c >> 1) & 1) * sy + (c & 1);
          if (valf(grid, q, sub, sign) > 0.0f) cas |= 1 << c;
        }
        cases[o++] = (unsigned char)cas;
      }
}
