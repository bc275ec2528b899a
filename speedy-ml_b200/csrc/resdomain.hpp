// resdomain.hpp -- region tiling of the 96x48x8 SPEEDY grid and the flattened gather/scatter maps.
//
// Host-side integer arithmetic only.  It produces, for every region, the same index sets as the
// reference's res_domain.f90 (cited per function, paths relative to the reference tree) but in closed
// form: halo columns are a modular walk, pole clipping is a min/max, and every pack/unpack routine is
// flattened once into an int32 offset list that the CUDA gather/scatter kernels consume.
// tests/test_index_maps.py checks the lists bit-for-bit against the oracle's slice-by-slice restatement.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace sml {

constexpr int XG = 96, YG = 48, ZG = 8, NVAR = 4;

// global device buffers (doubles):  G = [w4d | w2d | precip | sst | tisr],  F = [f4d | f2d]
constexpr int64_t G_W4D = 0;
constexpr int64_t G_W2D = (int64_t)NVAR * XG * YG * ZG;
constexpr int64_t G_PRECIP = G_W2D + XG * YG;
constexpr int64_t G_SST = G_PRECIP + XG * YG;
constexpr int64_t G_TISR = G_SST + XG * YG;
constexpr int64_t G_TOTAL = G_TISR + XG * YG;
constexpr int64_t F_F4D = 0;
constexpr int64_t F_F2D = G_W2D;
constexpr int64_t F_TOTAL = F_F2D + XG * YG;

// 0-based offsets, arguments 0-based
inline int64_t off4(int v, int x, int y, int z) { return v + (int64_t)NVAR * (x + (int64_t)XG * (y + (int64_t)YG * z)); }
inline int64_t off2(int x, int y) { return x + (int64_t)XG * y; }

struct Tiling {
    int fx = 0, fy = 0;  // grid points per region in x and y
    int ntx = 0, nty = 0;
    bool ok = false;
};

// domaindecomposition (src/res_domain.f90:258-280): the largest fy <= sqrt(4608/R) dividing 48 with
// fx = (4608/R)/fy dividing 96.  Region counts for which the Fortran loop would run into MOD(ygrid,0)
// are reported as !ok.
inline Tiling make_tiling(int num_regions)
{
    Tiling t;
    if (num_regions <= 0) return t;
    const int pts = (XG * YG) / num_regions;
    for (int fy = (int)std::floor(std::sqrt((double)pts)); fy >= 1; --fy) {
        if (YG % fy) continue;
        if (pts % fy) continue;
        const int fx = pts / fy;
        if (XG % fx) continue;
        t.fx = fx; t.fy = fy; t.ntx = XG / fx; t.nty = YG / fy; t.ok = true;
        return t;
    }
    return t;
}

struct RegionGeom {
    // 1-based inclusive, as res_domain.f90 reports them
    int xs, xe, ys, ye;          // interior (getxyresextent :123-141)
    int ixs, ixe, iys, iye;      // halo window (getoverlapindices :155-204)
    int ixc, iyc;                // halo extents
    int tdx0, tdy0;              // 0-based position of the interior inside the halo block
    bool pole, periodic;
    std::vector<int> gx;         // 0-based global x of each local halo column (wraps in x)
};

// getworkerlower_leftcorner (:282-292): regions run south->north fastest, then west->east.
inline RegionGeom make_geom(const Tiling &t, int region, int ov)
{
    RegionGeom g;
    const int cy = region % t.nty, cx = region / t.nty;
    g.xs = cx * t.fx + 1; g.xe = (cx + 1) * t.fx;
    g.ys = cy * t.fy + 1; g.ye = (cy + 1) * t.fy;
    g.periodic = (g.xs - ov < 1) || (g.xe + ov > XG);
    g.ixs = (g.xs - ov < 1) ? XG - ov + 1 : g.xs - ov;
    g.ixe = (g.xe + ov > XG) ? ov : g.xe + ov;
    g.ixc = t.fx + 2 * ov;
    g.iys = (g.ys - ov < 1) ? 1 : g.ys - ov;
    g.iye = (g.ye + ov > YG) ? YG : g.ye + ov;
    g.pole = (g.ys - ov < 1) || (g.ye + ov > YG);
    g.iyc = g.iye - g.iys + 1;
    g.gx.resize(g.ixc);
    // the two-slab periodic copy of tileoverlapgrid* (:380-418) is a modular walk from xs-ov
    for (int l = 0; l < g.ixc; ++l) g.gx[l] = (((g.xs - 1 - ov + l) % XG) + XG) % XG;
    g.tdx0 = ov;             // get_trainingdataindices (:547-574): x interior always starts after ov
    g.tdy0 = g.ys - g.iys;   // pole-clipped in y
    return g;
}

struct RegionSizes {
    int n, k, D, P, S, L, q;
    int logp_ms, tisr_ms, precip_ms, sst_ms;  // 0-based mean/std slots, -1 if absent
    int atmo_len, logp_off, precip_off, sst_off, tisr_off; // 0-based offsets in the input vector (-1 absent)
};

// allocate_res_new (src/mod_reservoir.f90:155-173) + trained_reservoir_prediction (:1822-1885), one vertical level
inline RegionSizes make_sizes(const Tiling &t, const RegionGeom &g, int m, double deg, bool precip,
                              bool sst_bool, bool sst_in, bool ml_only)
{
    RegionSizes s{};
    const int ixy = g.ixc * g.iyc, rxy = t.fx * t.fy;
    sst_in = sst_in && sst_bool;
    int L = NVAR * ZG;
    s.logp_ms = L++;
    s.tisr_ms = L++;
    s.precip_ms = precip ? L++ : -1;
    s.sst_ms = sst_bool ? L++ : -1;
    s.L = L;
    s.P = rxy * NVAR * ZG + rxy + (precip ? rxy : 0);
    s.S = ml_only ? 0 : rxy * NVAR * ZG + rxy;
    s.atmo_len = NVAR * ixy * ZG;
    s.logp_off = s.atmo_len;
    int nxt = s.logp_off + ixy;
    s.precip_off = -1; s.sst_off = -1;
    if (precip) { s.precip_off = nxt; nxt += ixy; }
    if (sst_in) { s.sst_off = nxt; nxt += ixy; }
    s.tisr_off = nxt; nxt += ixy;
    s.D = nxt;
    const double qd = (double)m / (double)s.D;
    s.q = (int)std::floor(qd + 0.5);       // NINT
    s.n = s.q * s.D;
    s.k = (int)((deg / (double)m) * s.n * s.n);  // real -> integer truncation (:172)
    return s;
}

struct RegionMaps {
    std::vector<int32_t> input_map, input_ms;    // [D] offsets into G / mean-std slot (sst slot -> L)
    std::vector<int32_t> output_map, output_ms;  // [P] offsets into G
    std::vector<int32_t> model_map, model_ms;    // [S] offsets into F
    std::vector<int32_t> target_map;             // [P] rows of the input vector
};

inline RegionMaps make_maps(const Tiling &t, const RegionGeom &g, const RegionSizes &s, bool precip, bool sst_in)
{
    RegionMaps m;
    m.input_map.assign(s.D, 0); m.input_ms.assign(s.D, -1);
    // atmosphere block (var, lx, ly, lz) -- tileoverlapgrid4d :348-420, standardised per (var, level)
    int e = 0;
    for (int lz = 0; lz < ZG; ++lz)
        for (int ly = 0; ly < g.iyc; ++ly)
            for (int lx = 0; lx < g.ixc; ++lx)
                for (int v = 0; v < NVAR; ++v) {
                    const int idx = v + NVAR * (lx + g.ixc * (ly + g.iyc * lz));
                    m.input_map[idx] = (int32_t)(G_W4D + off4(v, g.gx[lx], g.iys - 1 + ly, lz));
                    m.input_ms[idx] = v * ZG + lz;
                    ++e;
                }
    auto fill2d = [&](int base, int64_t goff, int ms) {
        for (int ly = 0; ly < g.iyc; ++ly)
            for (int lx = 0; lx < g.ixc; ++lx) {
                m.input_map[base + lx + g.ixc * ly] = (int32_t)(goff + off2(g.gx[lx], g.iys - 1 + ly));
                m.input_ms[base + lx + g.ixc * ly] = ms;
            }
    };
    fill2d(s.logp_off, G_W2D, s.logp_ms);
    if (precip) fill2d(s.precip_off, G_PRECIP, s.precip_ms);
    if (sst_in) fill2d(s.sst_off, G_SST, s.L);  // ocean reservoir's sst mean/std live in the extra slot L
    fill2d(s.tisr_off, G_TISR, s.tisr_ms);

    // outvec (var, rx, ry, z | logp | precip) -> global grids, tile_full_grid_with_local_state_vec_res1d :791-826
    m.output_map.assign(s.P, 0); m.output_ms.assign(s.P, -1);
    m.target_map.assign(s.P, 0);
    e = 0;
    for (int z = 0; z < ZG; ++z)
        for (int ry = 0; ry < t.fy; ++ry)
            for (int rx = 0; rx < t.fx; ++rx)
                for (int v = 0; v < NVAR; ++v) {
                    m.output_map[e] = (int32_t)(G_W4D + off4(v, g.xs - 1 + rx, g.ys - 1 + ry, z));
                    m.output_ms[e] = v * ZG + z;
                    m.target_map[e] = v + NVAR * ((g.tdx0 + rx) + g.ixc * ((g.tdy0 + ry) + g.iyc * z));
                    ++e;
                }
    for (int ry = 0; ry < t.fy; ++ry)
        for (int rx = 0; rx < t.fx; ++rx) {
            m.output_map[e] = (int32_t)(G_W2D + off2(g.xs - 1 + rx, g.ys - 1 + ry));
            m.output_ms[e] = s.logp_ms;
            m.target_map[e] = s.logp_off + (g.tdx0 + rx) + g.ixc * (g.tdy0 + ry);
            ++e;
        }
    if (precip)
        for (int ry = 0; ry < t.fy; ++ry)
            for (int rx = 0; rx < t.fx; ++rx) {
                m.output_map[e] = (int32_t)(G_PRECIP + off2(g.xs - 1 + rx, g.ys - 1 + ry));
                m.output_ms[e] = s.precip_ms;
                m.target_map[e] = s.precip_off + (g.tdx0 + rx) + g.ixc * (g.tdy0 + ry);
                ++e;
            }
    // local_model: same walk without precip, out of F (tile_4d_and_logp_full_grid_to_local_res_vec :1022-1053,
    // standardize_state_vec_res :1270-1315)
    m.model_map.assign(s.S, 0); m.model_ms.assign(s.S, -1);
    if (s.S > 0) {
        e = 0;
        for (int z = 0; z < ZG; ++z)
            for (int ry = 0; ry < t.fy; ++ry)
                for (int rx = 0; rx < t.fx; ++rx)
                    for (int v = 0; v < NVAR; ++v) {
                        m.model_map[e] = (int32_t)(F_F4D + off4(v, g.xs - 1 + rx, g.ys - 1 + ry, z));
                        m.model_ms[e] = v * ZG + z;
                        ++e;
                    }
        for (int ry = 0; ry < t.fy; ++ry)
            for (int rx = 0; rx < t.fx; ++rx) {
                m.model_map[e] = (int32_t)(F_F2D + off2(g.xs - 1 + rx, g.ys - 1 + ry));
                m.model_ms[e] = s.logp_ms;
                ++e;
            }
    }
    return m;
}

// ---- vertical localisation (num_vert_levels in {1, 2, 4, 8}) ---------------------------------------------
// get_z_res_extent (src/res_domain.f90:143-153), getoverlapindices_vert (:206-256), get_trainingdataindices_vert
// (:576-600).  1-based inclusive like the reference; tdz0 is the 0-based position of the reservoir's own levels
// inside its halo'd input column.
struct VertGeom {
    int zs = 1, ze = ZG, zc = ZG;        // levels the reservoir predicts
    int izs = 1, ize = ZG, izc = ZG;     // levels it reads (vertical overlap, clipped at the top / bottom)
    bool top = true, bottom = true;
    int tdzs = 1, tdze = ZG;             // get_trainingdataindices_vert: position of zs..ze inside izs..ize (1-based)
    bool ok = false;
};

inline VertGeom make_vert(int num_vert_levels, int vert_level, int vert_overlap)
{
    VertGeom v;
    if (num_vert_levels < 1 || ZG % num_vert_levels != 0 || vert_level < 1 || vert_level > num_vert_levels || vert_overlap < 0)
        return v;
    v.zc = ZG / num_vert_levels;
    v.zs = (vert_level - 1) * v.zc + 1;
    v.ze = vert_level * v.zc;
    v.top = v.zs == 1;
    v.bottom = v.ze == ZG;
    if (v.zs - vert_overlap >= 1 && v.ze + vert_overlap <= ZG) {
        v.izs = v.zs - vert_overlap; v.ize = v.ze + vert_overlap; v.izc = v.zc + 2 * vert_overlap;
    } else if (v.zs - vert_overlap < 1) {
        v.izs = 1; v.ize = v.ze + vert_overlap; v.izc = v.zc + vert_overlap + (v.zs - 1);
    } else {
        v.izs = v.zs - vert_overlap; v.ize = ZG; v.izc = v.zc + vert_overlap + (ZG - v.ze);
    }
    // the reference's second branch does not clip input_zend when BOTH ends overflow (it prints no error either):
    // such layouts (vert_overlap >= the slab height next to both boundaries) are rejected here
    if (v.ize > ZG || v.izs < 1) return v;
    if (v.zs - vert_overlap < 1) { v.tdzs = 1 + (v.zs - 1); v.tdze = v.izc - vert_overlap; }
    else if (v.ze + vert_overlap > ZG) { v.tdzs = 1 + vert_overlap; v.tdze = v.izc - (ZG - v.ze); }
    else { v.tdzs = 1 + vert_overlap; v.tdze = v.izc - vert_overlap; }
    v.ok = true;
    return v;
}

// allocate_res_new + trained_reservoir_prediction for ONE vertical slab (src/mod_reservoir.f90:80-180, 1793-1885):
// only the bottom slab carries logp, precip and SST; every slab carries TISR.  Sizes and vector offsets as make_sizes.
inline RegionSizes make_sizes_vert(const Tiling &t, const RegionGeom &g, const VertGeom &v, int m, double deg, bool precip,
                                   bool sst_bool, bool sst_in, bool ml_only)
{
    RegionSizes s{};
    const int ixy = g.ixc * g.iyc, rxy = t.fx * t.fy;
    const bool logp = v.bottom;
    precip = precip && v.bottom;
    sst_bool = sst_bool && v.bottom;
    sst_in = sst_in && sst_bool;
    int L = NVAR * v.izc;
    s.logp_ms = logp ? L++ : -1;
    s.tisr_ms = L++;
    s.precip_ms = precip ? L++ : -1;
    s.sst_ms = sst_bool ? L++ : -1;
    s.L = L;
    s.P = rxy * NVAR * v.zc + (logp ? rxy : 0) + (precip ? rxy : 0);
    s.S = ml_only ? 0 : rxy * NVAR * v.zc + (logp ? rxy : 0);
    s.atmo_len = NVAR * ixy * v.izc;
    int nxt = s.atmo_len;
    s.logp_off = -1; s.precip_off = -1; s.sst_off = -1;
    if (logp) { s.logp_off = nxt; nxt += ixy; }
    if (precip) { s.precip_off = nxt; nxt += ixy; }
    if (sst_in) { s.sst_off = nxt; nxt += ixy; }
    s.tisr_off = nxt; nxt += ixy;
    s.D = nxt;
    s.q = (int)std::floor((double)m / (double)s.D + 0.5);
    s.n = s.q * s.D;
    s.k = (int)((deg / (double)m) * s.n * s.n);
    return s;
}

// the flattened maps of one vertical slab: same walks as make_maps over the slab's own levels
inline RegionMaps make_maps_vert(const Tiling &t, const RegionGeom &g, const VertGeom &v, const RegionSizes &s, bool precip,
                                 bool sst_in)
{
    RegionMaps m;
    precip = precip && v.bottom;
    sst_in = sst_in && v.bottom;
    m.input_map.assign(s.D, 0); m.input_ms.assign(s.D, -1);
    for (int lz = 0; lz < v.izc; ++lz)
        for (int ly = 0; ly < g.iyc; ++ly)
            for (int lx = 0; lx < g.ixc; ++lx)
                for (int var = 0; var < NVAR; ++var) {
                    const int idx = var + NVAR * (lx + g.ixc * (ly + g.iyc * lz));
                    m.input_map[idx] = (int32_t)(G_W4D + off4(var, g.gx[lx], g.iys - 1 + ly, v.izs - 1 + lz));
                    m.input_ms[idx] = var * v.izc + lz;
                }
    auto fill2d = [&](int base, int64_t goff, int ms) {
        for (int ly = 0; ly < g.iyc; ++ly)
            for (int lx = 0; lx < g.ixc; ++lx) {
                m.input_map[base + lx + g.ixc * ly] = (int32_t)(goff + off2(g.gx[lx], g.iys - 1 + ly));
                m.input_ms[base + lx + g.ixc * ly] = ms;
            }
    };
    if (s.logp_off >= 0) fill2d(s.logp_off, G_W2D, s.logp_ms);
    if (precip) fill2d(s.precip_off, G_PRECIP, s.precip_ms);
    if (sst_in) fill2d(s.sst_off, G_SST, s.L);
    fill2d(s.tisr_off, G_TISR, s.tisr_ms);

    m.output_map.assign(s.P, 0); m.output_ms.assign(s.P, -1);
    m.target_map.assign(s.P, 0);
    int e = 0;
    for (int z = 0; z < v.zc; ++z)
        for (int ry = 0; ry < t.fy; ++ry)
            for (int rx = 0; rx < t.fx; ++rx)
                for (int var = 0; var < NVAR; ++var) {
                    m.output_map[e] = (int32_t)(G_W4D + off4(var, g.xs - 1 + rx, g.ys - 1 + ry, v.zs - 1 + z));
                    // unstandardize_state_vec_res (:1447-1457): the slot of input level tdata_zstart + z
                    m.output_ms[e] = var * v.izc + (v.tdzs - 1 + z);
                    m.target_map[e] = var + NVAR * ((g.tdx0 + rx) + g.ixc * ((g.tdy0 + ry) + g.iyc * (v.tdzs - 1 + z)));
                    ++e;
                }
    if (s.logp_off >= 0)
        for (int ry = 0; ry < t.fy; ++ry)
            for (int rx = 0; rx < t.fx; ++rx) {
                m.output_map[e] = (int32_t)(G_W2D + off2(g.xs - 1 + rx, g.ys - 1 + ry));
                m.output_ms[e] = s.logp_ms;
                m.target_map[e] = s.logp_off + (g.tdx0 + rx) + g.ixc * (g.tdy0 + ry);
                ++e;
            }
    if (precip)
        for (int ry = 0; ry < t.fy; ++ry)
            for (int rx = 0; rx < t.fx; ++rx) {
                m.output_map[e] = (int32_t)(G_PRECIP + off2(g.xs - 1 + rx, g.ys - 1 + ry));
                m.output_ms[e] = s.precip_ms;
                m.target_map[e] = s.precip_off + (g.tdx0 + rx) + g.ixc * (g.tdy0 + ry);
                ++e;
            }
    m.model_map.assign(s.S, 0); m.model_ms.assign(s.S, -1);
    if (s.S > 0) {
        e = 0;
        for (int z = 0; z < v.zc; ++z)
            for (int ry = 0; ry < t.fy; ++ry)
                for (int rx = 0; rx < t.fx; ++rx)
                    for (int var = 0; var < NVAR; ++var) {
                        m.model_map[e] = (int32_t)(F_F4D + off4(var, g.xs - 1 + rx, g.ys - 1 + ry, v.zs - 1 + z));
                        m.model_ms[e] = var * v.izc + (v.tdzs - 1 + z);
                        ++e;
                    }
        if (s.logp_off >= 0)
            for (int ry = 0; ry < t.fy; ++ry)
                for (int rx = 0; rx < t.fx; ++rx) {
                    m.model_map[e] = (int32_t)(F_F2D + off2(g.xs - 1 + rx, g.ys - 1 + ry));
                    m.model_ms[e] = s.logp_ms;
                    ++e;
                }
    }
    return m;
}

// ---- slab-ocean reservoir (res%reservoir_special / res%grid_special) ------------------------------------
// initialize_slab_ocean_model (src/mod_slab_ocean_reservoir.f90:9-133): input vector
//   [ atmosphere lowest level (var,lx,ly) 4*ixy | logp ixy | sst ixy | tisr ixy | ohtc ixy ],  D = 8*ixy,
//   output [ sst(rx,ry) | ohtc(rx,ry) ], P = 2*fx*fy, always ML-only (S = 0); offsets :1598-1618.
struct OceanSizes {
    int n, k, D, P, q;
    int ixy;       // halo tile points
    int A;         // logp_end: length of the time-averaged atmosphere part (5*ixy)
    int sst_off, tisr_off, ohtc_off;  // 0-based offsets in the ocean input vector
    int atmo_slice0;  // 0-based start of atmo_training_data_idx (:1621-1625) in the ATMOSPHERE input vector
};

inline OceanSizes make_ocean_sizes(const Tiling &t, const RegionGeom &g, int m, double deg)
{
    OceanSizes s{};
    s.ixy = g.ixc * g.iyc;
    s.A = NVAR * s.ixy + s.ixy;
    s.sst_off = s.A;
    s.tisr_off = s.sst_off + s.ixy;
    s.ohtc_off = s.tisr_off + s.ixy;
    s.D = s.ohtc_off + s.ixy;
    s.P = 2 * t.fx * t.fy;
    s.q = (int)std::floor((double)m / (double)s.D + 0.5);
    s.n = s.q * s.D;
    s.k = (int)((deg / (double)m) * s.n * s.n);
    s.atmo_slice0 = NVAR * s.ixy * ZG - NVAR * s.ixy;  // atmo3d_end - 4*ixy: the lowest level is the last z slab
    return s;
}

struct OceanMaps {
    std::vector<int32_t> sst_src;     // [ixy] offsets into G of the halo'd SST tile (tileoverlapgrid2d)
    std::vector<int32_t> target_map;  // [P] rows of the ocean input vector (tile_full_input_to_target_data2d_ocean_model, src/res_domain.f90:691-728)
};

inline OceanMaps make_ocean_maps(const Tiling &t, const RegionGeom &g, const OceanSizes &s)
{
    OceanMaps m;
    m.sst_src.assign(s.ixy, 0);
    for (int ly = 0; ly < g.iyc; ++ly)
        for (int lx = 0; lx < g.ixc; ++lx)
            m.sst_src[lx + g.ixc * ly] = (int32_t)(G_SST + off2(g.gx[lx], g.iys - 1 + ly));
    m.target_map.assign(s.P, 0);
    int e = 0;
    for (int blk = 0; blk < 2; ++blk) {
        const int base = blk == 0 ? s.sst_off : s.ohtc_off;
        for (int ry = 0; ry < t.fy; ++ry)
            for (int rx = 0; rx < t.fx; ++rx) m.target_map[e++] = base + (g.tdx0 + rx) + g.ixc * (g.tdy0 + ry);
    }
    return m;
}

// processor_decomposition (src/res_domain.f90:31-62)
inline std::vector<int32_t> regions_of_rank(int irank, int numprocs, int nregions)
{
    std::vector<int32_t> out;
    const int per = nregions / numprocs, left = nregions % numprocs;
    for (int i = 0; i < per; ++i) out.push_back(per * irank + i);
    if (irank >= 1 && irank <= left) out.push_back(nregions - left + irank - 1);
    return out;
}

}  // namespace sml
