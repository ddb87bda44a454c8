// field_kernels.cuh -- navigation-field construction on the device: everything
// Field.navigation_to_target (simulation/field.py:155-164) does between the geometry and the (U, V) maps the Navigation node
// samples:
//   distance_map            core/steering/quickest_path.py:54-117   raster of targets / obstacles + skfmm.distance
//   direction_map           quickest_path.py:144-163                normalised np.gradient (masked-array semantics)
//   fill_missing            quickest_path.py:168-181                nearest valid boundary value into the buffer zone
//   shortest_path           quickest_path.py:184-197                obstacles buffered by `radius` are impassable
//   direction_map_obstacles core/steering/obstacle_handling.py:106+ distance / direction from the walls
//   obstacle_handling       obstacle_handling.py:15-74              blend away from the walls within `radius`
// Third-party pieces of the reference that are absent here (skfmm 0.0.9, shapely buffer, skimage draw / find_boundaries) are
// replaced by their published algorithms: "parity unpinned" against those packages, pinned against oracle/field_oracle.*
// (same algorithms on the CPU) and against closed-form distances; direction_map / obstacle_handling follow the reference's
// own numpy / numba code, which the oracle's golden vectors pin.
//
// Eikonal solver: block-based fast iterative method.  The first-order upwind (Godunov) update of Sethian's fast marching --
//   T = min(a, b) + h  if |a - b| >= h,   (a + b + sqrt(2 h^2 - (a - b)^2)) / 2  otherwise --
// is iterated to its fixed point, which is the solution fast marching computes: 16 x 16 tiles are relaxed in shared memory,
// a tile whose values changed wakes its four neighbours for the next round, rounds run until no tile is active.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

constexpr int ET = 16;                    // tile edge
constexpr int EIK_SWEEPS = 2 * ET;        // relaxation sweeps per activation: enough for information to cross the tile

__device__ __forceinline__ double eik_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// quickest_path.py:41-44 indicer: ((p - min) / step).astype(int64)  (truncation toward zero)
__device__ __forceinline__ long long field_index(double p, double mn, double step) { return (long long)((p - mn) / step); }

// draw_geom for a LineString (core/geometry.py:112-116): skimage.draw.line between the indices of the end points (Bresenham)
__global__ void k_raster_lines(const double *__restrict__ seg, int n_seg, double minx, double miny, double step, int ny, int nx,
                               uint8_t *__restrict__ grid) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_seg) return;
    long long r = field_index(seg[4 * w], minx, step), c = field_index(seg[4 * w + 1], miny, step);
    const long long r1 = field_index(seg[4 * w + 2], minx, step), c1 = field_index(seg[4 * w + 3], miny, step);
    long long dr = llabs(r1 - r), dc = llabs(c1 - c);
    long long sr = (r1 - r) > 0 ? 1 : -1, sc = (c1 - c) > 0 ? 1 : -1;
    bool steep = false;
    if (dr > dc) { steep = true; long long t = c; c = r; r = t; t = dc; dc = dr; dr = t; t = sc; sc = sr; sr = t; }
    long long d = 2 * dr - dc;
    auto put = [&](long long x, long long y) { if (x >= 0 && x < nx && y >= 0 && y < ny) grid[y * nx + x] = 1; };
    for (long long i = 0; i < dc; ++i) {
        if (steep) put(c, r); else put(r, c);
        while (d >= 0) { r += sr; d -= 2 * dc; }
        c += sc; d += 2 * dr;
    }
    put(r1, c1);
}

// obstacles.buffer(radius) rasterised (quickest_path.py:186): cells whose grid point lies within `radius` of a segment
__global__ void k_buffer_mask(const double *__restrict__ seg, int n_seg, double radius, double minx, double miny, double step, int ny, int nx,
                              uint8_t *__restrict__ mask) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= (long long)ny * nx) return;
    const double x = minx + step * (double)(g % nx), y = miny + step * (double)(g / nx);
    bool in = false;
    for (int w = 0; w < n_seg && !in; ++w) {
        const double ax = seg[4 * w], ay = seg[4 * w + 1], bx = seg[4 * w + 2], by = seg[4 * w + 3];
        const double ex = bx - ax, ey = by - ay, l2 = ex * ex + ey * ey;
        double t = l2 > 0.0 ? ((x - ax) * ex + (y - ay) * ey) / l2 : 0.0;
        t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
        const double qx = x - (ax + t * ex), qy = y - (ay + t * ey);
        in = qx * qx + qy * qy <= radius * radius;
    }
    mask[g] = in ? 1 : 0;
}

// state: 0 free, 1 frozen (next to the zero level set), 2 masked
__global__ void k_eik_init(const uint8_t *__restrict__ target, const uint8_t *__restrict__ mask, int ny, int nx, double h, double *__restrict__ T,
                           uint8_t *__restrict__ state, uint8_t *__restrict__ active, int tiles_x, int tiles_y) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= (long long)ny * nx) return;
    const int x = (int)(g % nx), y = (int)(g / nx);
    if (mask && mask[g]) { state[g] = 2; T[g] = eik_inf(); return; }
    const int s = target[g] ? 1 : -1;
    auto opp = [&](int x2, int y2) {
        if (x2 < 0 || x2 >= nx || y2 < 0 || y2 >= ny) return false;
        const long long c = (long long)y2 * nx + x2;
        if (mask && mask[c]) return false;
        return (target[c] ? 1 : -1) != s;
    };
    const bool cx = opp(x - 1, y) || opp(x + 1, y), cy = opp(x, y - 1) || opp(x, y + 1);
    if (cx || cy) {
        const double d = h / 2.0;
        double inv = 0.0;
        if (cx) inv += 1.0 / (d * d);
        if (cy) inv += 1.0 / (d * d);
        T[g] = 1.0 / sqrt(inv);
        state[g] = 1;
        const int tx = x / ET, ty = y / ET;
        active[ty * tiles_x + tx] = 1;
        if (tx > 0) active[ty * tiles_x + tx - 1] = 1;
        if (tx + 1 < tiles_x) active[ty * tiles_x + tx + 1] = 1;
        if (ty > 0) active[(ty - 1) * tiles_x + tx] = 1;
        if (ty + 1 < tiles_y) active[(ty + 1) * tiles_x + tx] = 1;
    } else {
        T[g] = eik_inf();
        state[g] = 0;
    }
}

__device__ __forceinline__ double eik_update(double a, double b, double h) {
    const double lo = fmin(a, b), hi = fmax(a, b);
    if (isinf(lo)) return eik_inf();
    if (isinf(hi) || hi - lo >= h) return lo + h;
    return (a + b + sqrt(2.0 * h * h - (a - b) * (a - b))) / 2.0;
}

__global__ void __launch_bounds__(ET * ET) k_eik_fim(double *__restrict__ T, const uint8_t *__restrict__ state, int ny, int nx, double h,
                                                    const uint8_t *__restrict__ act_in, uint8_t *__restrict__ act_out, int tiles_x, int tiles_y,
                                                    unsigned *__restrict__ n_changed) {
    const int tile = blockIdx.y * tiles_x + blockIdx.x;
    if (!act_in[tile]) return;
    __shared__ double s[ET + 2][ET + 2];
    __shared__ int s_changed;
    const int lx = threadIdx.x, ly = threadIdx.y;
    const int x0 = blockIdx.x * ET, y0 = blockIdx.y * ET;
    if (lx == 0 && ly == 0) s_changed = 0;
    for (int k = ly * ET + lx; k < (ET + 2) * (ET + 2); k += ET * ET) {
        const int sx = k % (ET + 2), sy = k / (ET + 2);
        const int gx = x0 + sx - 1, gy = y0 + sy - 1;
        double v = eik_inf();
        if (gx >= 0 && gx < nx && gy >= 0 && gy < ny) v = T[(long long)gy * nx + gx];      // masked cells hold +inf
        s[sy][sx] = v;
    }
    const int gx = x0 + lx, gy = y0 + ly;
    const bool inside = gx < nx && gy < ny;
    const bool is_free = inside && state[(long long)gy * nx + gx] == 0;
    bool changed = false;
    __syncthreads();
    for (int it = 0; it < EIK_SWEEPS; ++it) {
        double t = s[ly + 1][lx + 1];
        if (is_free) {
            const double a = fmin(s[ly + 1][lx], s[ly + 1][lx + 2]), b = fmin(s[ly][lx + 1], s[ly + 2][lx + 1]);
            const double u = eik_update(a, b, h);
            if (u < t) { t = u; changed = true; }
        }
        __syncthreads();
        s[ly + 1][lx + 1] = t;
        __syncthreads();
    }
    if (is_free && changed) T[(long long)gy * nx + gx] = s[ly + 1][lx + 1];
    if (changed) s_changed = 1;
    __syncthreads();
    if (lx == 0 && ly == 0 && s_changed) {
        act_out[tile] = 1;
        if (blockIdx.x > 0) act_out[tile - 1] = 1;
        if ((int)blockIdx.x + 1 < tiles_x) act_out[tile + 1] = 1;
        if (blockIdx.y > 0) act_out[tile - tiles_x] = 1;
        if ((int)blockIdx.y + 1 < tiles_y) act_out[tile + tiles_x] = 1;
        atomicAdd(n_changed, 1u);
    }
}

// signed distance map like skfmm.distance: positive inside the raster, negative outside, NaN where masked
__global__ void k_eik_sign(double *__restrict__ T, const uint8_t *__restrict__ target, const uint8_t *__restrict__ state, long long n) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= n) return;
    T[g] = state[g] == 2 ? nan("") : (target[g] ? T[g] : -T[g]);
}

// np.gradient (unit spacing, edge_order 1) of one axis at index i of n, on values f(i); masked-array semantics: the result is
// masked where a value it uses is masked (the cell itself is not used by the central difference)
struct Grad { double g; bool masked; };
template <typename F, typename M>
__device__ __forceinline__ Grad grad_axis(int i, int n, F f, M m) {
    if (n == 1) return {0.0, m(0)};
    if (i == 0) return {f(1) - f(0), m(0) || m(1)};
    if (i == n - 1) return {f(n - 1) - f(n - 2), m(n - 1) || m(n - 2)};
    return {(f(i + 1) - f(i - 1)) / 2.0, m(i + 1) || m(i - 1)};
}

// direction_map (quickest_path.py:144-163): u, v = np.gradient(dmap); l = hypot(u, v); l[l == 0] = nan; (v / l, u / l)
// dirmask: 1 where the masked-array result is masked.  `cellmask` may be null (no masked cells: the obstacle map).
__device__ __forceinline__ void direction_at(const double *__restrict__ dmap, const uint8_t *__restrict__ cellmask, int x, int y, int ny, int nx,
                                             double &U, double &V, bool &masked) {
    auto fy = [&](int i) { return dmap[(long long)i * nx + x]; };
    auto my = [&](int i) { return cellmask && cellmask[(long long)i * nx + x] != 0; };
    auto fx = [&](int i) { return dmap[(long long)y * nx + i]; };
    auto mx = [&](int i) { return cellmask && cellmask[(long long)y * nx + i] != 0; };
    const Grad u = grad_axis(y, ny, fy, my), v = grad_axis(x, nx, fx, mx);     // axis 0 = rows (y), axis 1 = columns (x)
    masked = u.masked || v.masked;
    double l = hypot(u.g, v.g);
    if (l == 0.0) l = nan("");
    U = v.g / l; V = u.g / l;
}

__global__ void k_direction_map(const double *__restrict__ dmap, const uint8_t *__restrict__ cellmask, int ny, int nx, double *__restrict__ U,
                                double *__restrict__ V, uint8_t *__restrict__ dirmask) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= (long long)ny * nx) return;
    const int x = (int)(g % nx), y = (int)(g / nx);
    double u, v;
    bool m;
    direction_at(dmap, cellmask, x, y, ny, nx, u, v, m);
    U[g] = m ? nan("") : u; V[g] = m ? nan("") : v;
    dirmask[g] = m ? 1 : 0;
}

// fill_missing (quickest_path.py:168-181): cells that are masked in the direction map take the value of the NEAREST boundary
// cell -- an unmasked cell with a masked 4-neighbour (skimage find_boundaries(mode='outer') of the mask) -- by Euclidean
// distance between grid points (scipy NearestNDInterpolator).  Ties: the first candidate in row-major order (scipy's kd-tree
// does not define one).  Search window: `reach` cells; none found => stays NaN.
// The reference fills logical_xor(obstacle raster, mask), i.e. it leaves the cells ON the obstacle lines masked (undefined
// data that obstacle_handling then reads); here they are filled like the rest of the buffer zone, so that an agent pushed
// onto a wall line never samples an undefined direction.  Everywhere else the two rules select the same cells.
__global__ void k_fill_missing(const uint8_t *__restrict__ dirmask, const uint8_t *__restrict__ obstacle, int ny, int nx, int reach,
                               const double *__restrict__ U, const double *__restrict__ V, double *__restrict__ Uo, double *__restrict__ Vo) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= (long long)ny * nx) return;
    const int x = (int)(g % nx), y = (int)(g / nx);
    Uo[g] = U[g]; Vo[g] = V[g];
    (void)obstacle;
    if (dirmask[g] != 0) {
        long long best = -1;
        long long bd = 0x7fffffffffffffffLL;
        for (int y2 = max(y - reach, 0); y2 <= min(y + reach, ny - 1); ++y2)
            for (int x2 = max(x - reach, 0); x2 <= min(x + reach, nx - 1); ++x2) {
                const long long c = (long long)y2 * nx + x2;
                if (dirmask[c]) continue;
                const bool nb = (x2 > 0 && dirmask[c - 1]) || (x2 + 1 < nx && dirmask[c + 1]) || (y2 > 0 && dirmask[c - nx]) ||
                                (y2 + 1 < ny && dirmask[c + nx]);
                if (!nb) continue;
                const long long d = (long long)(x2 - x) * (x2 - x) + (long long)(y2 - y) * (y2 - y);
                if (d < bd) { bd = d; best = c; }
            }
        if (best >= 0) { Uo[g] = U[best]; Vo[g] = V[best]; }
    }
}

// obstacle_handling (obstacle_handling.py:15-74): within `radius` of the walls blend the direction away from them into the
// direction towards the target, p = strength ** (x / radius); then normalise everything.
__global__ void k_obstacle_handling(const double *__restrict__ dmap_obs, int ny, int nx, double radius, double strength, double *__restrict__ U,
                                    double *__restrict__ V) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= (long long)ny * nx) return;
    const int xi = (int)(g % nx), yi = (int)(g / nx);
    double u = U[g], v = V[g];
    const double x = -dmap_obs[g];
    if (0.0 < x && x < radius) {
        double u1, v1;
        bool m;
        direction_at(dmap_obs, nullptr, xi, yi, ny, nx, u1, v1, m);
        const double p = pow(strength, x / radius);
        u = -p * u1 + (1.0 - p) * u;
        v = -p * v1 + (1.0 - p) * v;
    }
    const double l = hypot(u, v);
    U[g] = u / l; V[g] = v / l;
}
