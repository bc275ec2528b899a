/*
 * speedyml_oracle.c -- CPU restatement (plain C, FP64) of the SPEEDY-ML reservoir hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see speedyml_oracle.h).  PARITY UNPINNED by the reference:
 * no Fortran/MPI/MKL toolchain exists here and the reference's tests hold no runnable
 * golden vector; fidelity rests on agreement with oracle/oracle_np.py and the two
 * known answers of tests/mod_unit_test.f90.
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 * The arithmetic follows the reference's ALGORITHMIC FORM on purpose (COO SpMV in entry
 * order, dense W_in GEMV over the stored n x D matrix, dense W_out GEMV, full-square
 * per-batch Gram, LU with partial pivoting) -- it is also the timed CPU baseline.
 * Compile with -ffp-contract=off so a*b+c is never fused where the reference rounds twice.
 */
#include "speedyml_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define XG ORC_XGRID
#define YG ORC_YGRID
#define ZG ORC_ZGRID

/* column-major helpers, 1-based arguments */
#define G4(nv, v, x, y, z) ((size_t)((v)-1) + (size_t)(nv) * ((size_t)((x)-1) + XG * ((size_t)((y)-1) + (size_t)YG * ((z)-1))))
#define G2(x, y) ((size_t)((x)-1) + (size_t)XG * ((y)-1))

/* ------------------------------------------------------------------ */
/* res_domain.f90 index arithmetic                                     */
/* ------------------------------------------------------------------ */

/* src/res_domain.f90:258-280 domaindecomposition.  The Fortran loop runs i=factorMax,0,-1 and
 * would evaluate MOD(ygrid,0) if it never exits; only region counts that exit are supported,
 * others return -1 here. */
int orc_domaindecomposition(int numregions, int *factorx, int *factory)
{
    if (numregions <= 0) return -1;
    int n = (XG * YG) / numregions; /* speedygridnum/numregions, integer division */
    int factorMax = (int)floor(sqrt((double)n));
    int fx = 0, fy = 0;
    for (int i = factorMax; i >= 1; --i) {
        if (YG % i == 0) {
            fy = i;
            if (n % fy == 0) {
                fx = n / fy;
                if (XG % fx == 0) {
                    *factorx = fx;
                    *factory = fy;
                    return 0;
                }
            }
        }
    }
    return -1;
}

/* src/res_domain.f90:282-292 getworkerlower_leftcorner */
void orc_getworkerlower_leftcorner(int region_num, int factory, int *row, int *col)
{
    *col = region_num % (YG / factory);
    *row = (int)floor((double)region_num / ((double)YG / (double)factory));
}

/* src/res_domain.f90:123-141 getxyresextent */
int orc_getxyresextent(int num_regions, int region_num, int *xs, int *xe, int *ys, int *ye,
                       int *xchunk, int *ychunk)
{
    int cornerx, cornery;
    if (orc_domaindecomposition(num_regions, xchunk, ychunk)) return -1;
    orc_getworkerlower_leftcorner(region_num, *ychunk, &cornerx, &cornery);
    *xs = cornerx * (*xchunk) + 1;
    *xe = (cornerx + 1) * (*xchunk);
    *ys = cornery * (*ychunk) + 1;
    *ye = (cornery + 1) * (*ychunk);
    return 0;
}

/* src/res_domain.f90:143-153 get_z_res_extent */
void orc_get_z_res_extent(int num_vert_levels, int vert_level, int *zs, int *ze, int *zchunk)
{
    *zchunk = ZG / num_vert_levels;
    *zs = (vert_level - 1) * (*zchunk) + 1;
    *ze = vert_level * (*zchunk);
}

/* src/res_domain.f90:155-204 getoverlapindices */
int orc_getoverlapindices(int numregions, int region_num, int overlap, int *ixs, int *ixe, int *iys,
                          int *iye, int *ixc, int *iyc, int *pole, int *periodic)
{
    int xs, xe, ys, ye, xc, yc;
    if (orc_getxyresextent(numregions, region_num, &xs, &xe, &ys, &ye, &xc, &yc)) return -1;
    *ixc = xc + 2 * overlap;
    *iyc = yc + 2 * overlap;
    *periodic = 0;
    *pole = 0;
    if (xs - overlap < 1) {
        *ixs = XG - overlap + 1;
        *periodic = 1;
    } else {
        *ixs = xs - overlap;
    }
    if (xe + overlap > XG) {
        *ixe = overlap;
        *periodic = 1;
    } else {
        *ixe = overlap + xe;
    }
    if (ys - overlap < 1) {
        *iys = 1;
        *iyc = yc + overlap + (ys - 1);
        *pole = 1;
    } else {
        *iys = ys - overlap;
    }
    if (ye + overlap > YG) {
        *iye = YG;
        *iyc = yc + overlap + (YG - ye);
        *pole = 1;
    } else {
        *iye = overlap + ye;
    }
    return 0;
}

/* src/res_domain.f90:206-256 getoverlapindices_vert */
int orc_getoverlapindices_vert(int num_vert_levels, int vert_level, int vert_overlap, int *izs,
                               int *ize, int *izc, int *top, int *bottom)
{
    int zs, ze, zc;
    orc_get_z_res_extent(num_vert_levels, vert_level, &zs, &ze, &zc);
    *top = (zs == 1);
    *bottom = (ze == ZG);
    if (zs - vert_overlap >= 1 && ze + vert_overlap <= ZG) {
        *izs = zs - vert_overlap;
        *ize = ze + vert_overlap;
        *izc = zc + 2 * vert_overlap;
    } else if (zs - vert_overlap < 1) {
        *izs = 1;
        *ize = ze + vert_overlap;
        *izc = zc + vert_overlap + (zs - 1);
    } else if (ze + vert_overlap > ZG) {
        *izs = zs - vert_overlap;
        *ize = ZG;
        *izc = zc + vert_overlap + (ZG - ze);
    } else {
        return -1; /* 'something is wrong' branch */
    }
    return 0;
}

/* src/res_domain.f90:547-574 get_trainingdataindices */
int orc_get_trainingdataindices(int num_regions, int region_num, int overlap, int *xs, int *xe,
                                int *ys, int *ye)
{
    int rxs, rxe, rys, rye, rxc, ryc, ixs, ixe, iys, iye, ixc, iyc, pole, per;
    if (orc_getxyresextent(num_regions, region_num, &rxs, &rxe, &rys, &rye, &rxc, &ryc)) return -1;
    orc_getoverlapindices(num_regions, region_num, overlap, &ixs, &ixe, &iys, &iye, &ixc, &iyc, &pole, &per);
    *xs = 1 + overlap;
    *xe = ixc - overlap;
    if (rys - overlap < 1) {
        *ys = 1 + (rys - 1);
        *ye = iyc - overlap;
    } else if (rye + overlap > YG) {
        *ys = 1 + overlap;
        *ye = iyc - (YG - rye);
    } else {
        *ys = 1 + overlap;
        *ye = iyc - overlap;
    }
    return 0;
}

/* src/res_domain.f90:576-600 get_trainingdataindices_vert */
void orc_get_trainingdataindices_vert(int num_vert_levels, int vert_level, int vert_overlap, int *zs,
                                      int *ze)
{
    int rzs, rze, rzc, izs, ize, izc, top, bottom;
    orc_get_z_res_extent(num_vert_levels, vert_level, &rzs, &rze, &rzc);
    orc_getoverlapindices_vert(num_vert_levels, vert_level, vert_overlap, &izs, &ize, &izc, &top, &bottom);
    if (rzs - vert_overlap < 1) {
        *zs = 1 + (rzs - 1);
        *ze = izc - vert_overlap;
    } else if (rze + vert_overlap > ZG) {
        *zs = 1 + vert_overlap;
        *ze = izc - (ZG - rze);
    } else {
        *zs = 1 + vert_overlap;
        *ze = izc - vert_overlap;
    }
}

/* src/res_domain.f90:31-62 processor_decomposition (twin :64-94).  Returns the count. */
int orc_processor_decomposition(int irank, int numprocs, int number_of_regions, int *region_indices)
{
    int per = number_of_regions / numprocs;
    int left_over = number_of_regions % numprocs;
    int i;
    if (irank >= left_over + 1 && irank > 0) {
        for (i = 1; i <= per; ++i) region_indices[i - 1] = per * irank + i - 1;
        return per;
    } else if (irank == 0) {
        for (i = 1; i <= per; ++i) region_indices[i - 1] = i - 1;
        return per;
    } else {
        for (i = 1; i <= per; ++i) region_indices[i - 1] = per * irank + i - 1;
        region_indices[per] = number_of_regions - left_over + irank - 1;
        return per + 1;
    }
}

/* src/mod_utilities.f90:1598-1636 find_closest_divisor */
int orc_find_closest_divisor(int target, int number)
{
    if (((number % target) + target) % target == 0) return target;
    int radius = 2;
    for (;;) {
        for (int i = target - radius; i <= target + radius; ++i) {
            if (i != 0 && number % i == 0) return i;
        }
        radius++;
    }
}

/* NINT: round half away from zero */
static int nint_d(double v) { return (int)(v >= 0.0 ? floor(v + 0.5) : -floor(-v + 0.5)); }

/* initializedomain (src/res_domain.f90:96-121) + flag setup and offsets of
 * trained_reservoir_prediction (src/mod_reservoir.f90:1783-1886) + sizes of allocate_res_new
 * (src/mod_reservoir.f90:80-180). */
int orc_setup_region(int num_regions, int region, int overlap, int num_vert_levels, int vert_level,
                     int vert_overlap, int m, double deg, int precip_bool, int slab_ocean_model_bool,
                     int sst_bool_input, int ml_only, orc_grid *g, orc_dims *d)
{
    memset(g, 0, sizeof(*g));
    memset(d, 0, sizeof(*d));
    g->number_of_regions = num_regions;
    g->region = region;
    g->overlap = overlap;
    g->num_vert_levels = num_vert_levels;
    g->level_index = vert_level;
    g->vert_overlap = vert_overlap;
    if (orc_getxyresextent(num_regions, region, &g->res_xstart, &g->res_xend, &g->res_ystart,
                           &g->res_yend, &g->resxchunk, &g->resychunk))
        return -1;
    orc_get_z_res_extent(num_vert_levels, vert_level, &g->res_zstart, &g->res_zend, &g->reszchunk);
    orc_getoverlapindices(num_regions, region, overlap, &g->input_xstart, &g->input_xend,
                          &g->input_ystart, &g->input_yend, &g->inputxchunk, &g->inputychunk,
                          &g->pole, &g->periodicboundary);
    if (orc_getoverlapindices_vert(num_vert_levels, vert_level, vert_overlap, &g->input_zstart,
                                   &g->input_zend, &g->inputzchunk, &g->top, &g->bottom))
        return -1;
    orc_get_trainingdataindices(num_regions, region, overlap, &g->tdata_xstart, &g->tdata_xend,
                                &g->tdata_ystart, &g->tdata_yend);
    orc_get_trainingdataindices_vert(num_vert_levels, vert_level, vert_overlap, &g->tdata_zstart,
                                     &g->tdata_zend);

    /* src/mod_reservoir.f90:1793-1818 */
    if (g->bottom) {
        d->logp_bool = 1;
        d->tisr_input_bool = 1;
        d->sst_bool = slab_ocean_model_bool;
        d->precip_input_bool = precip_bool;
        d->precip_bool = precip_bool;
    } else {
        d->logp_bool = 0;
        d->tisr_input_bool = 1;
        d->sst_bool = 0;
        d->precip_input_bool = 0;
        d->precip_bool = 0;
    }
    d->local_predictvars = 4;
    d->local_heightlevels_input = g->inputzchunk;
    d->local_heightlevels_res = g->reszchunk;
    d->ml_only = ml_only;

    /* src/mod_reservoir.f90:1822-1848 mean/std slots */
    int msl = 4 * g->inputzchunk;
    if (d->logp_bool) g->logp_mean_std_idx = ++msl;
    if (d->tisr_input_bool) g->tisr_mean_std_idx = ++msl;
    if (d->precip_bool) g->precip_mean_std_idx = ++msl;
    if (d->sst_bool) {
        g->sst_mean_std_idx = ++msl;
        d->sst_bool_input = sst_bool_input ? 1 : 0;
    }
    g->mean_std_length = msl;

    /* src/mod_reservoir.f90:91-173 allocate_res_new */
    d->m = m;
    d->deg = deg;
    d->leakage = 1.0;
    d->density = deg / (double)m;
    int ixy = g->inputxchunk * g->inputychunk, rxy = g->resxchunk * g->resychunk;
    d->logp_size_input = d->logp_bool ? ixy : 0;
    d->sst_size_input = d->sst_bool_input ? ixy : 0;
    d->logp_size_res = d->logp_bool ? rxy : 0;
    d->precip_size_res = d->precip_input_bool ? rxy : 0;
    d->precip_size_input = d->precip_input_bool ? ixy : 0;
    d->sst_size_res = d->sst_bool_input ? rxy : 0;
    d->tisr_size_res = d->tisr_input_bool ? rxy : 0;
    d->tisr_size_input = d->tisr_input_bool ? ixy : 0;
    d->chunk_size = rxy * d->local_predictvars * g->reszchunk + d->logp_size_res + d->precip_size_res;
    d->chunk_size_prediction = d->chunk_size;
    d->chunk_size_speedy = rxy * d->local_predictvars * g->reszchunk + d->logp_size_res;
    if (ml_only) d->chunk_size_speedy = 0;
    d->locality = ixy * g->inputzchunk * d->local_predictvars + d->logp_size_input +
                  d->precip_size_input + d->tisr_size_input + d->sst_size_input - d->chunk_size;
    d->nodes_per_input = nint_d((double)d->m / ((double)d->chunk_size + (double)d->locality));
    d->n = d->nodes_per_input * (d->chunk_size + d->locality);
    d->k = (int)(d->density * d->n * d->n); /* real -> integer assignment truncates, :172 */
    d->reservoir_numinputs = d->chunk_size + d->locality;

    /* src/mod_reservoir.f90:1854-1885 vector offsets */
    g->atmo3d_start = 1;
    g->atmo3d_end = 4 * ixy * g->inputzchunk;
    g->predict_start = 1;
    g->predict_end = g->atmo3d_end;
    if (d->logp_bool) {
        g->logp_start = g->atmo3d_end + 1;
        g->logp_end = g->atmo3d_end + d->logp_size_input;
        g->predict_end = g->logp_end;
    }
    if (d->precip_bool) {
        g->precip_start = g->atmo3d_end + d->logp_size_input + 1;
        g->precip_end = g->precip_start + d->precip_size_input - 1;
        g->predict_end = g->precip_end;
    }
    if (d->sst_bool_input) {
        g->sst_start = g->atmo3d_end + d->logp_size_input + d->precip_size_input + 1;
        g->sst_end = g->sst_start + d->sst_size_input - 1;
    }
    if (d->tisr_input_bool) {
        g->tisr_start = g->atmo3d_end + d->logp_size_input + d->precip_size_input + d->sst_size_input + 1;
        g->tisr_end = g->tisr_start + d->tisr_size_input - 1;
    }
    return 0;
}

/* slab-ocean reservoir (res%reservoir_special / res%grid_special).
 * sizes: initialize_slab_ocean_model, src/mod_slab_ocean_reservoir.f90:9-133 (m=4000, deg=6, leakage=1,
 * always ML-only :26,112-114); vector offsets: trained_ocean_reservoir_prediction :1598-1640.
 * grid_special starts as a copy of the bottom-level atmosphere grid (same tiling, same mean/std slots). */
int orc_setup_ocean_region(int num_regions, int region, int overlap, int m, double deg, int precip_bool,
                           orc_grid *g, orc_dims *d)
{
    orc_dims da;
    if (orc_setup_region(num_regions, region, overlap, 1, 1, 0, 6000, 6.0, precip_bool, 1, 1, 0, g, &da)) return -1;
    memset(d, 0, sizeof(*d));
    const int ixy = g->inputxchunk * g->inputychunk, rxy = g->resxchunk * g->resychunk;
    d->local_predictvars = 4;                      /* :18 full_predictvars */
    d->local_heightlevels_input = g->inputzchunk;  /* :19 */
    d->local_heightlevels_res = g->reszchunk;      /* :21 */
    d->ml_only = 1;                                /* :23 ml_only_ocean */
    d->m = m;
    d->deg = deg;
    d->density = deg / (double)m;                  /* :37 */
    d->leakage = 1.0;                              /* :41 */
    d->sst_bool = 1;
    d->sst_bool_input = 1;
    d->tisr_input_bool = 1;
    d->sst_size_res = rxy;                         /* :57 */
    d->sst_size_input = ixy;                       /* :63 */
    d->tisr_size_res = rxy;                        /* :75 */
    d->tisr_size_input = ixy;                      /* :81 */
    const int atmo_size_input = ixy * d->local_predictvars + ixy; /* :87 (precip_input_bool is .False. :1597) */
    const int ohtc_input_size = ixy, ohtc_res_size = rxy;         /* :98-108 */
    d->chunk_size_speedy = 0;                      /* :112-114 */
    d->chunk_size = d->sst_size_res + ohtc_res_size;            /* :116 */
    d->chunk_size_prediction = d->sst_size_res + ohtc_res_size; /* :118 */
    d->locality = atmo_size_input + d->sst_size_input + 0 + d->tisr_size_input + ohtc_input_size - d->chunk_size; /* :122 */
    d->nodes_per_input = nint_d((double)d->m / ((double)d->chunk_size + (double)d->locality));
    d->n = d->nodes_per_input * (d->chunk_size + d->locality);
    d->k = (int)(d->density * d->n * d->n);
    d->reservoir_numinputs = d->chunk_size + d->locality;
    /* :1598-1618 */
    g->atmo3d_start = 1;
    g->atmo3d_end = ixy * 4;
    g->logp_start = g->atmo3d_end + 1;
    g->logp_end = ixy * 4 + ixy;
    g->precip_start = g->precip_end = 0;
    g->sst_start = g->logp_end + 1;
    g->sst_end = g->sst_start + ixy - 1;
    g->tisr_start = g->sst_end + 1;
    g->tisr_end = g->tisr_start + ixy - 1;
    g->ohtc_start = g->tisr_end + 1;
    g->ohtc_end = g->ohtc_start + ixy - 1;
    g->ohtc_mean_std_idx = 1; /* :1640 */
    g->is_ocean = 1;
    return 0;
}

/* ------------------------------------------------------------------ */
/* tilers                                                              */
/* ------------------------------------------------------------------ */

/* local x index (1-based) -> global x, restating the two-slab copy of tileoverlapgrid*
 * (src/res_domain.f90:380-418 and the 2d/3d/5d twins). */
static int local_to_global_x(int lx, int ixs, int ixe, int periodic, int rxs, int rxe)
{
    if (periodic && (rxe > ixe || ixs > rxs)) {
        int first = XG - (ixs - 1); /* local 1..first <- global ixs..XG */
        if (lx <= first) return ixs + lx - 1;
        return lx - first; /* local first+1.. <- global 1..ixe */
    }
    return ixs + lx - 1;
}

/* src/res_domain.f90:348-420 tileoverlapgrid4d: grid4d(nvars,96,48,8) -> localgrid(nvars,ixc,iyc,izc) */
int orc_tileoverlapgrid4d(const double *grid4d, int nvars, int numregions, int region, int overlap,
                          int num_vert_levels, int vert_level, int vert_overlap, double *localgrid)
{
    int rxs, rxe, rys, rye, rxc, ryc, ixs, ixe, iys, iye, ixc, iyc, pole, per;
    int izs, ize, izc, top, bottom;
    if (orc_getxyresextent(numregions, region, &rxs, &rxe, &rys, &rye, &rxc, &ryc)) return -1;
    orc_getoverlapindices(numregions, region, overlap, &ixs, &ixe, &iys, &iye, &ixc, &iyc, &pole, &per);
    if (orc_getoverlapindices_vert(num_vert_levels, vert_level, vert_overlap, &izs, &ize, &izc, &top, &bottom))
        return -1;
    for (int lz = 1; lz <= izc; ++lz)
        for (int ly = 1; ly <= iyc; ++ly)
            for (int lx = 1; lx <= ixc; ++lx) {
                int gx = local_to_global_x(lx, ixs, ixe, per, rxs, rxe);
                int gy = iys + ly - 1, gz = izs + lz - 1;
                for (int v = 1; v <= nvars; ++v)
                    localgrid[(size_t)(v - 1) + (size_t)nvars * ((lx - 1) + (size_t)ixc * ((ly - 1) + (size_t)iyc * (lz - 1)))] =
                        grid4d[G4(nvars, v, gx, gy, gz)];
            }
    return 0;
}

/* src/res_domain.f90:484-545 tileoverlapgrid2d */
int orc_tileoverlapgrid2d(const double *grid2d, int numregions, int region, int overlap, double *localgrid)
{
    int rxs, rxe, rys, rye, rxc, ryc, ixs, ixe, iys, iye, ixc, iyc, pole, per;
    if (orc_getxyresextent(numregions, region, &rxs, &rxe, &rys, &rye, &rxc, &ryc)) return -1;
    orc_getoverlapindices(numregions, region, overlap, &ixs, &ixe, &iys, &iye, &ixc, &iyc, &pole, &per);
    for (int ly = 1; ly <= iyc; ++ly)
        for (int lx = 1; lx <= ixc; ++lx) {
            int gx = local_to_global_x(lx, ixs, ixe, per, rxs, rxe);
            localgrid[(lx - 1) + (size_t)ixc * (ly - 1)] = grid2d[G2(gx, iys + ly - 1)];
        }
    return 0;
}

/* src/res_domain.f90:1081-1125 tile_4d_and_logp_to_local_state_input */
int orc_tile_4d_and_logp_to_local_state_input(int numregions, int region, int overlap, int num_vert_levels,
                                              int vert_level, int vert_overlap, int precip_bool,
                                              const double *grid4d, const double *grid2d,
                                              const double *precip_grid, double *inputvec)
{
    int rzs, rze, rzc, ixs, ixe, iys, iye, ixc, iyc, pole, per, izs, ize, izc, top, bottom;
    orc_get_z_res_extent(num_vert_levels, vert_level, &rzs, &rze, &rzc);
    if (orc_getoverlapindices(numregions, region, overlap, &ixs, &ixe, &iys, &iye, &ixc, &iyc, &pole, &per)) return -1;
    orc_getoverlapindices_vert(num_vert_levels, vert_level, vert_overlap, &izs, &ize, &izc, &top, &bottom);
    int nv = 4;
    size_t n4 = (size_t)nv * ixc * iyc * izc, n2 = (size_t)ixc * iyc;
    orc_tileoverlapgrid4d(grid4d, nv, numregions, region, overlap, num_vert_levels, vert_level, vert_overlap, inputvec);
    if (rze == ZG) {
        orc_tileoverlapgrid2d(grid2d, numregions, region, overlap, inputvec + n4);
        if (precip_bool) orc_tileoverlapgrid2d(precip_grid, numregions, region, overlap, inputvec + n4 + n2);
    }
    return 0;
}

/* src/res_domain.f90:791-826 tile_full_grid_with_local_state_vec_res1d */
void orc_tile_full_grid_with_local_state_vec_res1d(int numregions, int region, int num_vert_levels,
                                                   int vert_level, int precip_bool, const double *statevec,
                                                   int length, double *wholegrid4d, double *wholegrid2d,
                                                   double *wholegrid_precip)
{
    int xs, xe, ys, ye, xc, yc, zs, ze, zc;
    (void)length;
    orc_getxyresextent(numregions, region, &xs, &xe, &ys, &ye, &xc, &yc);
    orc_get_z_res_extent(num_vert_levels, vert_level, &zs, &ze, &zc);
    const int nv = 4;
    size_t e = 0;
    for (int z = zs; z <= ze; ++z)
        for (int y = ys; y <= ye; ++y)
            for (int x = xs; x <= xe; ++x)
                for (int v = 1; v <= nv; ++v) wholegrid4d[G4(nv, v, x, y, z)] = statevec[e++];
    if (ze == ZG) {
        for (int y = ys; y <= ye; ++y)
            for (int x = xs; x <= xe; ++x) wholegrid2d[G2(x, y)] = statevec[e++];
        if (precip_bool)
            for (int y = ys; y <= ye; ++y)
                for (int x = xs; x <= xe; ++x) wholegrid_precip[G2(x, y)] = statevec[e++];
    } else {
        for (int y = ys; y <= ye; ++y)
            for (int x = xs; x <= xe; ++x) wholegrid2d[G2(x, y)] = 0.0;
    }
}

/* src/res_domain.f90:828-850 tile_full_2d_grid_with_local_res (first xc*yc entries reshaped) */
void orc_tile_full_2d_grid_with_local_res(int numregions, int region, const double *statevec, double *wholegrid2d)
{
    int xs, xe, ys, ye, xc, yc;
    orc_getxyresextent(numregions, region, &xs, &xe, &ys, &ye, &xc, &yc);
    size_t e = 0;
    for (int y = ys; y <= ye; ++y)
        for (int x = xs; x <= xe; ++x) wholegrid2d[G2(x, y)] = statevec[e++];
}

/* src/res_domain.f90:1022-1053 tile_4d_and_logp_full_grid_to_local_res_vec */
void orc_tile_4d_and_logp_full_grid_to_local_res_vec(int numregions, int region, int num_vert_levels,
                                                     int vert_level, const double *grid4d,
                                                     const double *grid2d, double *statevec)
{
    int xs, xe, ys, ye, xc, yc, zs, ze, zc;
    orc_getxyresextent(numregions, region, &xs, &xe, &ys, &ye, &xc, &yc);
    orc_get_z_res_extent(num_vert_levels, vert_level, &zs, &ze, &zc);
    const int nv = 4;
    size_t e = 0;
    for (int z = zs; z <= ze; ++z)
        for (int y = ys; y <= ye; ++y)
            for (int x = xs; x <= xe; ++x)
                for (int v = 1; v <= nv; ++v) statevec[e++] = grid4d[G4(nv, v, x, y, z)];
    if (ze == ZG)
        for (int y = ys; y <= ye; ++y)
            for (int x = xs; x <= xe; ++x) statevec[e++] = grid2d[G2(x, y)];
}

/* src/res_domain.f90:602-651 tile_full_input_to_target_data2d: statevec(ld, ncols) -> tiled(P, ncols) */
void orc_tile_full_input_to_target_data2d(const orc_grid *g, const orc_dims *d, const double *statevec,
                                          int ld, int ncols, double *tiled)
{
    const int nv = d->local_predictvars, ixc = g->inputxchunk, iyc = g->inputychunk;
    const int P = d->chunk_size_prediction;
    const int n4res = nv * g->resxchunk * g->resychunk * d->local_heightlevels_res;
    const int rxy = g->resxchunk * g->resychunk;
    if (g->is_ocean) {
        /* tile_full_input_to_target_data2d_ocean_model, src/res_domain.f90:691-728 (ohtc_prediction branch):
         * interior of the SST block, then interior of the OHTC block */
        for (int c = 0; c < ncols; ++c) {
            const double *sv = statevec + (size_t)c * ld;
            double *t = tiled + (size_t)c * P;
            size_t e = 0;
            const double *ss = sv + (g->sst_start - 1), *oh = sv + (g->ohtc_start - 1);
            for (int y = g->tdata_ystart; y <= g->tdata_yend; ++y)
                for (int x = g->tdata_xstart; x <= g->tdata_xend; ++x) t[e++] = ss[(x - 1) + (size_t)ixc * (y - 1)];
            for (int y = g->tdata_ystart; y <= g->tdata_yend; ++y)
                for (int x = g->tdata_xstart; x <= g->tdata_xend; ++x) t[e++] = oh[(x - 1) + (size_t)ixc * (y - 1)];
        }
        return;
    }
    for (int c = 0; c < ncols; ++c) {
        const double *sv = statevec + (size_t)c * ld;
        double *t = tiled + (size_t)c * P;
        size_t e = 0;
        for (int z = g->tdata_zstart; z <= g->tdata_zend; ++z)
            for (int y = g->tdata_ystart; y <= g->tdata_yend; ++y)
                for (int x = g->tdata_xstart; x <= g->tdata_xend; ++x)
                    for (int v = 1; v <= nv; ++v)
                        t[e++] = sv[(size_t)(v - 1) + (size_t)nv * ((x - 1) + (size_t)ixc * ((y - 1) + (size_t)iyc * (z - 1)))];
        if (d->logp_bool) {
            const double *lp = sv + (g->logp_start - 1);
            e = n4res;
            for (int y = g->tdata_ystart; y <= g->tdata_yend; ++y)
                for (int x = g->tdata_xstart; x <= g->tdata_xend; ++x) t[e++] = lp[(x - 1) + (size_t)ixc * (y - 1)];
        }
        if (d->precip_bool) {
            const double *pp = sv + (g->precip_start - 1);
            e = n4res + rxy;
            for (int y = g->tdata_ystart; y <= g->tdata_yend; ++y)
                for (int x = g->tdata_xstart; x <= g->tdata_xend; ++x) t[e++] = pp[(x - 1) + (size_t)ixc * (y - 1)];
        }
    }
}

/* src/res_domain.f90:1211-1268 standardize_state_vec_input (+ input_grid_to_input_statevec_and_standardization).
 * standardize_data_given_pars2d is (x - mean) then / std, two roundings (src/mod_utilities.f90:1307-1317). */
void orc_standardize_state_vec_input(const orc_grid *g, const orc_dims *d, const double *mean,
                                     const double *std, double *state_vec)
{
    const int nv = d->local_predictvars, ixc = g->inputxchunk, iyc = g->inputychunk, izc = g->inputzchunk;
    int l = 1;
    for (int i = 1; i <= nv; ++i)
        for (int j = 1; j <= izc; ++j) {
            for (int y = 1; y <= iyc; ++y)
                for (int x = 1; x <= ixc; ++x) {
                    size_t e = (size_t)(i - 1) + (size_t)nv * ((x - 1) + (size_t)ixc * ((y - 1) + (size_t)iyc * (j - 1)));
                    double v = state_vec[e] - mean[l - 1];
                    state_vec[e] = v / std[l - 1];
                }
            l++;
        }
    if (d->logp_bool) {
        double *lp = state_vec + (g->logp_start - 1);
        for (int e = 0; e < ixc * iyc; ++e) {
            double v = lp[e] - mean[l - 1];
            lp[e] = v / std[l - 1];
        }
    }
}

/* src/res_domain.f90:1270-1315 standardize_state_vec_res (local_model, length chunk_size_speedy) */
void orc_standardize_state_vec_res(const orc_grid *g, const orc_dims *d, const double *mean,
                                   const double *std, double *state_vec)
{
    const int nv = d->local_predictvars, rxc = g->resxchunk, ryc = g->resychunk;
    const int height = d->local_heightlevels_input;
    int l = 1;
    for (int i = 1; i <= nv; ++i) {
        int data_height = 1;
        for (int j = 1; j <= height; ++j) {
            if (j >= g->tdata_zstart && j <= g->tdata_zend) {
                for (int y = 1; y <= ryc; ++y)
                    for (int x = 1; x <= rxc; ++x) {
                        size_t e = (size_t)(i - 1) + (size_t)nv * ((x - 1) + (size_t)rxc * ((y - 1) + (size_t)ryc * (data_height - 1)));
                        double v = state_vec[e] - mean[l - 1];
                        state_vec[e] = v / std[l - 1];
                    }
                data_height++;
            }
            l++;
        }
    }
    if (d->logp_bool) {
        double *lp = state_vec + (size_t)nv * rxc * ryc * d->local_heightlevels_res;
        for (int e = 0; e < rxc * ryc; ++e) {
            double v = lp[e] - mean[l - 1];
            lp[e] = v / std[l - 1];
        }
    }
}

/* src/res_domain.f90:1424-1475 unstandardize_state_vec_res; unstandardize_data_2d is x*std then +mean
 * (two roundings, src/mod_utilities.f90:799-829). */
void orc_unstandardize_state_vec_res(const orc_grid *g, const orc_dims *d, const double *mean,
                                     const double *std, double *state_vec)
{
    const int nv = d->local_predictvars, rxc = g->resxchunk, ryc = g->resychunk;
    const int height = d->local_heightlevels_input;
    const size_t n4 = (size_t)nv * rxc * ryc * d->local_heightlevels_res;
    int l = 1;
    for (int i = 1; i <= nv; ++i) {
        int data_height = 1;
        for (int j = 1; j <= height; ++j) {
            if (j >= g->tdata_zstart && j <= g->tdata_zend) {
                for (int y = 1; y <= ryc; ++y)
                    for (int x = 1; x <= rxc; ++x) {
                        size_t e = (size_t)(i - 1) + (size_t)nv * ((x - 1) + (size_t)rxc * ((y - 1) + (size_t)ryc * (data_height - 1)));
                        double v = state_vec[e] * std[l - 1];
                        state_vec[e] = v + mean[l - 1];
                    }
                data_height++;
            }
            l++;
        }
    }
    if (d->logp_bool) {
        const int li = g->logp_mean_std_idx - 1;
        for (int e = 0; e < rxc * ryc; ++e) {
            double v = state_vec[n4 + e] * std[li];
            state_vec[n4 + e] = v + mean[li];
        }
    }
    if (d->precip_bool) {
        const int pi = g->precip_mean_std_idx - 1;
        for (int e = 0; e < rxc * ryc; ++e) {
            double v = state_vec[n4 + rxc * ryc + e] * std[pi];
            state_vec[n4 + rxc * ryc + e] = v + mean[pi];
        }
    }
}

/* ------------------------------------------------------------------ */
/* mod_linalg.f90                                                      */
/* ------------------------------------------------------------------ */

/* ------------------------------------------------------------------ */
/* reservoir construction: makesparse + shuffle + the W_in build        */
/* ------------------------------------------------------------------ */

/* The reference draws from the Fortran random_number stream, which cannot be reproduced.  The
 * engine and this restatement use a COUNTER-BASED generator instead: the draw for (seed, region, stream, index) is a
 * pure function of those four numbers (splitmix64 finaliser, the one behind the training noise), so the two sides can
 * be compared bit for bit whatever order the draws are made in.  Streams: 0 = vals, 1 + 2*round = row shuffle of a
 * round, 2 + 2*round = column shuffle, ORC_STREAM_WIN = the W_in values. */
static unsigned long long orc_mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
unsigned long long orc_counter_bits(unsigned long long seed, int region, int stream, long long index)
{
    const unsigned long long h1 = orc_mix64(seed ^ ((unsigned long long)(unsigned)region << 32) ^ (unsigned long long)(unsigned)stream);
    return orc_mix64(h1 + (unsigned long long)index);
}
/* random_number into a default real (shuffle's `real :: a`): 24 random bits, [0, 1) */
static float orc_u24(unsigned long long bits) { return (float)(bits >> 40) * (1.0f / 16777216.0f); }
/* random_number into real(dp): 53 random bits, [0, 1) */
static double orc_u53(unsigned long long bits) { return (double)(bits >> 11) * (1.0 / 9007199254740992.0); }

/* src/mod_utilities.f90:1569-1596 shuffle(n, returnsize, shufflereturn): the "k-shuffle".  choices = 1..n; step i
 * picks this = a*(n - n_chosen) + 1 in DEFAULT-REAL arithmetic (a is `real`, the product and the sum are single
 * precision, the assignment to the integer truncates), swaps it to the end of the live range.  Only the first
 * returnsize picks are returned, so the loop may stop there.  Guard (not in the reference): `this` is clamped to the
 * live range -- for n - n_chosen = 8192 and the largest a the single-precision product rounds up to 8192 and the
 * reference would read one element past the live range. */
void orc_shuffle(int n, int returnsize, unsigned long long seed, int region, int stream, int *shufflereturn)
{
    int *choices = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; ++i) choices[i] = i + 1;
    int n_chosen = 0;
    for (int i = 0; i < returnsize; ++i) {
        const float a = orc_u24(orc_counter_bits(seed, region, stream, i));
        const int live = n - n_chosen;
        const float prod = a * (float)live;
        const float sum = prod + 1.0f;
        int pick = (int)sum;
        if (pick > live) pick = live;
        const int tmp = choices[pick - 1];
        shufflereturn[i] = tmp;
        choices[pick - 1] = choices[live - 1];
        choices[live - 1] = tmp;
        n_chosen++;
    }
    free(choices);
}

/* src/mod_linalg.f90:180-218 makesparse: vals = random_number; rows / cols = `counter` full k-shuffles each plus a
 * partial one of `leftover` picks (k > n), or one partial shuffle each (k <= n).  1-based indices. */
void orc_makesparse(int n, int k, unsigned long long seed, int region, int *rows, int *cols, double *vals)
{
    for (int e = 0; e < k; ++e) vals[e] = orc_u53(orc_counter_bits(seed, region, 0, e));
    if (k > n) {
        const int counter = k / n, leftover = k % n;
        int i;
        for (i = 0; i < counter; ++i) {
            orc_shuffle(n, n, seed, region, 1 + 2 * i, rows + (size_t)i * n);
            orc_shuffle(n, n, seed, region, 2 + 2 * i, cols + (size_t)i * n);
        }
        if (leftover != 0) {
            orc_shuffle(n, leftover, seed, region, 1 + 2 * i, rows + (size_t)i * n);
            orc_shuffle(n, leftover, seed, region, 2 + 2 * i, cols + (size_t)i * n);
        }
    } else {
        orc_shuffle(n, k, seed, region, 1, rows);
        orc_shuffle(n, k, seed, region, 2, cols);
    }
}

/* src/mod_reservoir.f90:262-283: q = n / reservoir_numinputs; win = 0; for every input i the rows (i-1)q+1 .. iq get
 * sigma * (-1 + 2*rand).  Returned in the one-per-row form: winc[j] and the 0-based column wcol[j] = j / q. */
void orc_gen_win(int n, int D, double sigma, unsigned long long seed, int region, double *winc, int *wcol)
{
    const int q = n / D;
    for (int j = 0; j < n; ++j) { winc[j] = 0.0; wcol[j] = 0; }
    for (int i = 0; i < D; ++i)
        for (int t = 0; t < q; ++t) {
            const int j = i * q + t;
            const double rnd = orc_u53(orc_counter_bits(seed, region, ORC_STREAM_WIN, j));
            const double ip = -1.0 + 2.0 * rnd;
            winc[j] = sigma * ip;
            wcol[j] = i;
        }
}

/* MKL_SPARSE_D_MV on the COO handle of mklsparse (src/mod_linalg.f90:10-25): y = A*x, 1-based
 * indices, general matrix, duplicate (row,col) entries sum.  alpha=1, beta=0. */
void orc_coo_mv(int n, int k, const int *rows, const int *cols, const double *vals, const double *x, double *y)
{
    for (int i = 0; i < n; ++i) y[i] = 0.0;
    for (int e = 0; e < k; ++e) y[rows[e] - 1] += vals[e] * x[cols[e] - 1];
}

/* LAPACK dgesv restated: LU with partial pivoting (right-looking, column-major) then the two
 * triangular solves.  Returns info (0 ok, i>0: U(i,i) exactly zero).  src/mod_linalg.f90:145. */
int orc_dgesv(int n, int nrhs, double *A, int lda, int *ipiv, double *B, int ldb)
{
    int info = 0;
    const int NB = 48;
    for (int j0 = 0; j0 < n; j0 += NB) {
        int jb = (n - j0 < NB) ? n - j0 : NB;
        /* panel factorisation (dgetf2) on A[j0:n, j0:j0+jb] */
        for (int j = j0; j < j0 + jb; ++j) {
            int p = j;
            double amax = fabs(A[j + (size_t)lda * j]);
            for (int i = j + 1; i < n; ++i) {
                double v = fabs(A[i + (size_t)lda * j]);
                if (v > amax) { amax = v; p = i; }
            }
            ipiv[j] = p + 1;
            if (A[p + (size_t)lda * j] != 0.0) {
                if (p != j)
                    for (int c = 0; c < n; ++c) {
                        double t = A[j + (size_t)lda * c];
                        A[j + (size_t)lda * c] = A[p + (size_t)lda * c];
                        A[p + (size_t)lda * c] = t;
                    }
                double piv = 1.0 / A[j + (size_t)lda * j];
                for (int i = j + 1; i < n; ++i) A[i + (size_t)lda * j] *= piv;
            } else if (info == 0) {
                info = j + 1;
            }
            /* update the rest of the panel */
            for (int c = j + 1; c < j0 + jb; ++c) {
                double f = A[j + (size_t)lda * c];
                if (f != 0.0)
                    for (int i = j + 1; i < n; ++i) A[i + (size_t)lda * c] -= A[i + (size_t)lda * j] * f;
            }
        }
        int j1 = j0 + jb;
        if (j1 < n) {
            /* U12 = L11^-1 A12 */
#pragma omp parallel for schedule(static)
            for (int c = j1; c < n; ++c)
                for (int j = j0; j < j1; ++j) {
                    double f = A[j + (size_t)lda * c];
                    if (f != 0.0)
                        for (int i = j + 1; i < j1; ++i) A[i + (size_t)lda * c] -= A[i + (size_t)lda * j] * f;
                }
            /* A22 -= L21 U12 */
#pragma omp parallel for schedule(static)
            for (int c = j1; c < n; ++c)
                for (int j = j0; j < j1; ++j) {
                    double f = A[j + (size_t)lda * c];
                    if (f != 0.0)
                        for (int i = j1; i < n; ++i) A[i + (size_t)lda * c] -= A[i + (size_t)lda * j] * f;
                }
        }
    }
    if (info != 0) return info;
    /* dgetrs: apply row swaps to B, solve L then U */
    for (int j = 0; j < n; ++j) {
        int p = ipiv[j] - 1;
        if (p != j)
            for (int c = 0; c < nrhs; ++c) {
                double t = B[j + (size_t)ldb * c];
                B[j + (size_t)ldb * c] = B[p + (size_t)ldb * c];
                B[p + (size_t)ldb * c] = t;
            }
    }
#pragma omp parallel for schedule(static)
    for (int c = 0; c < nrhs; ++c) {
        double *b = B + (size_t)ldb * c;
        for (int j = 0; j < n; ++j) {
            double f = b[j];
            if (f != 0.0)
                for (int i = j + 1; i < n; ++i) b[i] -= f * A[i + (size_t)lda * j];
        }
        for (int j = n - 1; j >= 0; --j) {
            if (b[j] != 0.0) {
                b[j] /= A[j + (size_t)lda * j];
                double f = b[j];
                for (int i = 0; i < j; ++i) b[i] -= f * A[i + (size_t)lda * j];
            }
        }
    }
    return 0;
}

/* src/mod_linalg.f90:109-151 mldivide: A(n,m), B(l,k); returns -1 (A,B unchanged) when n != l,
 * otherwise dgesv's info (B holds the solution only when info == 0; reference prints and continues). */
int orc_mldivide(double *A, int n, int m, double *B, int l, int k)
{
    (void)m;
    if (n != l) return -1;
    int *ipiv = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int info = orc_dgesv(n, k, A, n > 1 ? n : 1, ipiv, B, n > 1 ? n : 1);
    free(ipiv);
    return info;
}

/* ------------------------------------------------------------------ */
/* reservoir object                                                    */
/* ------------------------------------------------------------------ */
struct orc_region {
    orc_grid g;
    orc_dims d;
    int *rows, *cols;
    double *vals, *win, *wout, *mean, *std;
    double *x, *feedback, *local_model, *outvec;
    /* optional one-non-zero-per-row form of win (test scaffolding for full-model runs; bit-identical to
     * the dense GEMV because every other term of matmul(win,u) is an exact zero -- tests/test_oracle_predict.py) */
    double *winc;
    int *wcol;
    /* training */
    int batch_size;
    double *states, *augmented_states, *saved_state, *sxs, *sxt; /* states_x_states_aug, states_x_trainingdata_aug */
};

orc_region *orc_region_new(const orc_grid *g, const orc_dims *d)
{
    orc_region *r = (orc_region *)calloc(1, sizeof(orc_region));
    r->g = *g;
    r->d = *d;
    const int n = d->n, D = d->reservoir_numinputs, P = d->chunk_size_prediction, S = d->chunk_size_speedy;
    r->rows = (int *)calloc((size_t)d->k + 1, sizeof(int));
    r->cols = (int *)calloc((size_t)d->k + 1, sizeof(int));
    r->vals = (double *)calloc((size_t)d->k + 1, sizeof(double));
    r->win = NULL; /* allocated by orc_region_set_weights when a dense win is given */
    r->wout = (double *)calloc((size_t)P * (n + S), sizeof(double));
    r->mean = (double *)calloc((size_t)g->mean_std_length + 1, sizeof(double));
    r->std = (double *)calloc((size_t)g->mean_std_length + 1, sizeof(double));
    r->x = (double *)calloc((size_t)n, sizeof(double));
    r->feedback = (double *)calloc((size_t)D, sizeof(double));
    r->local_model = (double *)calloc((size_t)S + 1, sizeof(double));
    r->outvec = (double *)calloc((size_t)P, sizeof(double));
    return r;
}

void orc_train_free(orc_region *r)
{
    free(r->states); free(r->augmented_states); free(r->saved_state); free(r->sxs); free(r->sxt);
    r->states = r->augmented_states = r->saved_state = r->sxs = r->sxt = NULL;
}

void orc_region_free(orc_region *r)
{
    if (!r) return;
    orc_train_free(r);
    free(r->rows); free(r->cols); free(r->vals); free(r->win); free(r->wout); free(r->mean); free(r->std);
    free(r->x); free(r->feedback); free(r->local_model); free(r->outvec);
    free(r->winc); free(r->wcol);
    free(r);
}

/* what read_trained_res delivers (src/mod_io.f90:2938-2983): win(n,D) dense, wout(P,n+S),
 * rows/cols(k) 1-based, vals(k), mean(L), std(L) */
int orc_region_set_weights(orc_region *r, const int *rows, const int *cols, const double *vals,
                           const double *win, const double *wout, const double *mean, const double *std,
                           int mean_std_length)
{
    const orc_dims *d = &r->d;
    const int n = d->n, D = d->reservoir_numinputs, P = d->chunk_size_prediction, S = d->chunk_size_speedy;
    if (mean_std_length != r->g.mean_std_length) return -1;
    for (int e = 0; e < d->k; ++e)
        if (rows[e] < 1 || rows[e] > n || cols[e] < 1 || cols[e] > n) return -2; /* mklsparse would stop */
    memcpy(r->rows, rows, sizeof(int) * (size_t)d->k);
    memcpy(r->cols, cols, sizeof(int) * (size_t)d->k);
    memcpy(r->vals, vals, sizeof(double) * (size_t)d->k);
    if (win) {
        if (!r->win) r->win = (double *)malloc(sizeof(double) * (size_t)n * D);
        memcpy(r->win, win, sizeof(double) * (size_t)n * D);
        free(r->winc); free(r->wcol);
        r->winc = NULL; r->wcol = NULL;
    }
    if (wout) memcpy(r->wout, wout, sizeof(double) * (size_t)P * (n + S));
    memcpy(r->mean, mean, sizeof(double) * (size_t)mean_std_length);
    memcpy(r->std, std, sizeof(double) * (size_t)mean_std_length);
    return 0;
}

void orc_region_set_leakage(orc_region *r, double leakage) { r->d.leakage = leakage; }

/* compact W_in: value and 0-based column of the single non-zero of each row (src/mod_reservoir.f90:270-280).
 * Frees the dense copy. */
/* expand the one-per-row form back into the dense win(n, D) the reference stores (src/mod_reservoir.f90:262-283) and
 * drop the compact copy: predict then runs the reference's dense matmul(win, feedback) again.  Used by bench.py's CPU
 * arm, which must time the reference's algorithmic form for all 1152 regions without pushing 30 GB of zeros through
 * Python first. */
int orc_region_densify_win(orc_region *r)
{
    if (!r->winc) return r->win ? 0 : 1;
    const int n = r->d.n, D = r->d.reservoir_numinputs;
    /* malloc + memset, not calloc: the reference stores real zeros (reservoir%win = 0.0_dp, then the non-zeros), so
     * its GEMV streams 8*n*D bytes from memory; untouched calloc pages would all alias the kernel's zero page */
    double *w = (double *)malloc((size_t)n * D * sizeof(double));
    if (!w) return 2;
    memset(w, 0, (size_t)n * D * sizeof(double));
    for (int j = 0; j < n; ++j) w[(size_t)r->wcol[j] * n + j] = r->winc[j];
    free(r->win);
    r->win = w;
    free(r->winc); free(r->wcol);
    r->winc = NULL; r->wcol = NULL;
    return 0;
}

int orc_region_set_win_compact(orc_region *r, const double *winc, const int *wcol)
{
    const int n = r->d.n, D = r->d.reservoir_numinputs;
    for (int j = 0; j < n; ++j)
        if (wcol[j] < 0 || wcol[j] >= D) return -1;
    free(r->winc); free(r->wcol);
    r->winc = (double *)malloc(sizeof(double) * (size_t)n);
    r->wcol = (int *)malloc(sizeof(int) * (size_t)n);
    memcpy(r->winc, winc, sizeof(double) * (size_t)n);
    memcpy(r->wcol, wcol, sizeof(int) * (size_t)n);
    free(r->win);
    r->win = NULL;
    return 0;
}
const orc_grid *orc_region_grid(const orc_region *r) { return &r->g; }
const orc_dims *orc_region_dims(const orc_region *r) { return &r->d; }

double *orc_region_ptr(orc_region *r, const char *f)
{
    if (!strcmp(f, "x")) return r->x;
    if (!strcmp(f, "feedback")) return r->feedback;
    if (!strcmp(f, "local_model")) return r->local_model;
    if (!strcmp(f, "outvec")) return r->outvec;
    if (!strcmp(f, "wout")) return r->wout;
    if (!strcmp(f, "win")) return r->win;
    if (!strcmp(f, "vals")) return r->vals;
    if (!strcmp(f, "mean")) return r->mean;
    if (!strcmp(f, "std")) return r->std;
    if (!strcmp(f, "saved_state")) return r->saved_state;
    if (!strcmp(f, "states_x_states_aug")) return r->sxs;
    if (!strcmp(f, "states_x_trainingdata_aug")) return r->sxt;
    return NULL;
}

/* temp = matmul(win, u): dense n x D GEMV in axpy (column) order, as the stored matrix dictates */
static void dense_win_gemv(const orc_region *r, const double *u, double *temp)
{
    const int n = r->d.n, D = r->d.reservoir_numinputs;
    if (r->winc) {
        for (int j = 0; j < n; ++j) temp[j] = r->winc[j] * u[r->wcol[j]];
        return;
    }
    for (int j = 0; j < n; ++j) temp[j] = 0.0;
    for (int i = 0; i < D; ++i) {
        const double ui = u[i];
        const double *col = r->win + (size_t)i * n;
        for (int j = 0; j < n; ++j) temp[j] += col[j] * ui;
    }
}

/* one state update: y = A x; temp = W_in u; x = (1-leak) x + leak tanh(y+temp)
 * src/mod_reservoir.f90:1373-1377, 1444-1448 */
static void state_update(const orc_region *r, const double *u, double *x, double *y, double *temp)
{
    const int n = r->d.n;
    const double leak = r->d.leakage;
    orc_coo_mv(n, r->d.k, r->rows, r->cols, r->vals, x, y);
    dense_win_gemv(r, u, temp);
    for (int j = 0; j < n; ++j) {
        double x_ = tanh(y[j] + temp[j]);
        double a = (1.0 - leak) * x[j];
        double b = leak * x_;
        x[j] = a + b;
    }
}

/* src/mod_reservoir.f90:1354-1381 synchronize (synchronize_print :1383-1416 is the same arithmetic);
 * slab twin src/mod_slab_ocean_reservoir.f90:1237-1266.  input is (ld, length) column-major. */
void orc_synchronize(orc_region *r, const double *input, int ld, double *x, int length)
{
    const int n = r->d.n;
    double *y = (double *)malloc(sizeof(double) * (size_t)n), *temp = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < length; ++i) state_update(r, input + (size_t)i * ld, x, y, temp);
    free(y);
    free(temp);
}

/* readout: x_temp = x; x_temp(2:n:2) squared; x_aug = [local_model ; x_temp]; outvec = matmul(wout, x_aug)
 * src/mod_reservoir.f90:1450-1456 */
static void readout(orc_region *r, const double *x, int S)
{
    const int n = r->d.n, P = r->d.chunk_size_prediction;
    const int N = n + S;
    double *xa = (double *)malloc(sizeof(double) * (size_t)N);
    for (int i = 0; i < S; ++i) xa[i] = r->local_model[i];
    for (int j = 0; j < n; ++j) xa[S + j] = ((j + 1) % 2 == 0) ? x[j] * x[j] : x[j];
    for (int p = 0; p < P; ++p) r->outvec[p] = 0.0;
    for (int j = 0; j < N; ++j) {
        const double xj = xa[j];
        const double *col = r->wout + (size_t)j * P;
        for (int p = 0; p < P; ++p) r->outvec[p] += col[p] * xj;
    }
    free(xa);
}

/* src/mod_reservoir.f90:1418-1489 predict */
void orc_predict(orc_region *r, double *x)
{
    const int n = r->d.n;
    double *y = (double *)malloc(sizeof(double) * (size_t)n), *temp = (double *)malloc(sizeof(double) * (size_t)n);
    state_update(r, r->feedback, x, y, temp);
    readout(r, x, r->d.chunk_size_speedy);
    orc_unstandardize_state_vec_res(&r->g, &r->d, r->mean, r->std, r->outvec);
    free(y);
    free(temp);
}

/* src/mod_reservoir.f90:1491-1535 predict_ml (chunk_size_speedy == 0) */
void orc_predict_ml(orc_region *r, double *x)
{
    const int n = r->d.n;
    double *y = (double *)malloc(sizeof(double) * (size_t)n), *temp = (double *)malloc(sizeof(double) * (size_t)n);
    state_update(r, r->feedback, x, y, temp);
    readout(r, x, 0);
    orc_unstandardize_state_vec_res(&r->g, &r->d, r->mean, r->std, r->outvec);
    free(y);
    free(temp);
}

/* the region loop of src/parallelmain.f90:226-251, one OpenMP thread per "rank" */
void orc_predict_all(orc_region **regs, int nreg, int ml_only, int nthreads)
{
    (void)nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int i = 0; i < nreg; ++i) {
        if (ml_only) orc_predict_ml(regs[i], regs[i]->x);
        else orc_predict(regs[i], regs[i]->x);
    }
}

/* src/mod_slab_ocean_reservoir.f90:1318-1363 predict_slab_ml: same update and readout (S = 0), then
 * outvec*std(sst)+mean(sst) on every output (SST and OHTC alike, :1354) */
void orc_predict_slab_ml(orc_region *r, double *x)
{
    const int n = r->d.n, P = r->d.chunk_size_prediction;
    double *y = (double *)malloc(sizeof(double) * (size_t)n), *temp = (double *)malloc(sizeof(double) * (size_t)n);
    state_update(r, r->feedback, x, y, temp);
    readout(r, x, 0);
    const int si = r->g.sst_mean_std_idx - 1;
    for (int p = 0; p < P; ++p) {
        double v = r->outvec[p] * r->std[si];
        r->outvec[p] = v + r->mean[si];
    }
    free(y);
    free(temp);
}

/* src/mod_slab_ocean_reservoir.f90:1268-1316 predict_slab (the HYBRID ocean reservoir, ml_only_ocean = .False.):
 * x = tanh(A x + W_in u) with no leak term (:1294), features [local_model ; x~] with local_model of length
 * chunk_size_prediction (:1299-1300), outvec = wout * features, then the quirk local_model = outvec BEFORE the
 * un-standardisation (:1303) -- the reservoir's own standardised prediction is its "imperfect model" of the next
 * step -- and outvec*std(sst)+mean(sst).  The region must have been created with chunk_size_speedy =
 * chunk_size_prediction. */
void orc_predict_slab(orc_region *r, double *x)
{
    const int n = r->d.n, P = r->d.chunk_size_prediction;
    double *y = (double *)malloc(sizeof(double) * (size_t)n), *temp = (double *)malloc(sizeof(double) * (size_t)n);
    const double leak = r->d.leakage;
    r->d.leakage = 1.0;                       /* x = tanh(y + temp): (1-1)*x + 1*tanh(.) is exactly tanh(.) */
    state_update(r, r->feedback, x, y, temp);
    r->d.leakage = leak;
    readout(r, x, P);
    const int si = r->g.sst_mean_std_idx - 1;
    for (int p = 0; p < P; ++p) {
        r->local_model[p] = r->outvec[p];
        double v = r->outvec[p] * r->std[si];
        r->outvec[p] = v + r->mean[si];
    }
    free(y);
    free(temp);
}

/* ocean feedback, intended semantics: src/mpires.f90:594-600 (SST tile, standardised
 * with grid_special's SST mean/std), :776-781 (ring of the atmosphere reservoir's standardised lowest-level
 * + logp feedback, slot mod(timestep-1,nslots)+1, mean = sum/nslots even while slots are still zero). */
void orc_ocean_feedback(orc_region *ocean, const orc_region *atmo, double *ring, int nslots, int timestep,
                        const double *wholegrid_sst)
{
    const orc_grid *g = &ocean->g, *ga = &atmo->g;
    const int ixy = g->inputxchunk * g->inputychunk;
    const int A = g->logp_end;
    const int a0 = ga->atmo3d_end - ixy * 4; /* 0-based start of atmo_training_data_idx (:1621-1625) */
    const int slot = (timestep - 1) % nslots;
    for (int e = 0; e < A; ++e) ring[e + (size_t)A * slot] = atmo->feedback[a0 + e];
    for (int e = 0; e < A; ++e) {
        double s = 0.0;
        for (int k = 0; k < nslots; ++k) s += ring[e + (size_t)A * k];
        ocean->feedback[e] = s / (double)nslots;
    }
    double *sst = ocean->feedback + (g->sst_start - 1);
    orc_tileoverlapgrid2d(wholegrid_sst, g->number_of_regions, g->region, g->overlap, sst);
    const int si = g->sst_mean_std_idx - 1;
    for (int e = 0; e < ixy; ++e) {
        double v = sst[e] - ocean->mean[si];
        sst[e] = v / ocean->std[si];
    }
}

/* ------------------------------------------------------------------ */
/* sendrecievegrid (src/mpires.f90:218-804), split at the host-model call */
/* ------------------------------------------------------------------ */

/* root's assembly :281-331 (+ the receive loop :401-454, same tiler) and the clamps :456-490 */
void orc_step_gather(orc_region **regs, int nreg, int precip_bool, int ocean_model,
                     const double *base_sst_grid, const double *sea_mask, const double *ocean_outvec,
                     const int *has_ocean, double *wholegrid4d, double *wholegrid2d,
                     double *wholegrid_precip, double *wholegrid_sst)
{
    const size_t n4 = (size_t)4 * XG * YG * ZG, n2 = (size_t)XG * YG;
    for (size_t e = 0; e < n4; ++e) wholegrid4d[e] = 0.0;
    for (size_t e = 0; e < n2; ++e) wholegrid2d[e] = 0.0;
    if (precip_bool)
        for (size_t e = 0; e < n2; ++e) wholegrid_precip[e] = 0.0;
    if (ocean_model)
        for (size_t e = 0; e < n2; ++e) wholegrid_sst[e] = base_sst_grid[e];
    for (int i = 0; i < nreg; ++i) {
        orc_region *r = regs[i];
        orc_tile_full_grid_with_local_state_vec_res1d(r->g.number_of_regions, r->g.region, r->g.num_vert_levels,
                                                      r->g.level_index, precip_bool, r->outvec,
                                                      r->d.chunk_size_prediction, wholegrid4d, wholegrid2d,
                                                      wholegrid_precip);
        if (ocean_model) {
            const int rxy = r->g.resxchunk * r->g.resychunk;
            double tmp[64];
            if (has_ocean && has_ocean[i] && ocean_outvec) {
                orc_tile_full_2d_grid_with_local_res(r->g.number_of_regions, r->g.region,
                                                     ocean_outvec + (size_t)i * rxy, wholegrid_sst);
            } else {
                for (int e = 0; e < rxy && e < 64; ++e) tmp[e] = 272.0; /* :323-326 */
                orc_tile_full_2d_grid_with_local_res(r->g.number_of_regions, r->g.region, tmp, wholegrid_sst);
            }
        }
    }
    /* :460-462 */
    for (size_t p = 0; p < (size_t)XG * YG * ZG; ++p)
        if (wholegrid4d[3 + 4 * p] < 0.000001) wholegrid4d[3 + 4 * p] = 0.000001;
    if (ocean_model) {
        for (size_t e = 0; e < n2; ++e) /* :470-478 */
            if (sea_mask[e] > 0.0) wholegrid_sst[e] = base_sst_grid[e];
        for (size_t e = 0; e < n2; ++e) /* :480-484 */
            if (wholegrid_sst[e] < 272.0) wholegrid_sst[e] = 272.0;
    }
    if (precip_bool)
        for (size_t e = 0; e < n2; ++e) /* :486-490 */
            if (wholegrid_precip[e] < 0.00001) wholegrid_precip[e] = 0.0;
}

/* the q floor applied to SPEEDY's input copy and output (src/mpires.f90:1583-1585, 1648-1650) */
void orc_run_model_clamp(double *grid4d)
{
    for (size_t p = 0; p < (size_t)XG * YG * ZG; ++p)
        if (grid4d[3 + 4 * p] < 0.000001) grid4d[3 + 4 * p] = 0.000001;
}

/* Deterministic stand-in for agcm_main: forecast = 0.98*grid + 0.02*climatology.
 * Not part of the reference; used identically by the CPU baseline and the GPU bench. */
void orc_host_stub(const double *grid4d, const double *grid2d, const double *clim4d, const double *clim2d,
                   double *forecast_4d, double *forecast_2d)
{
    const size_t n4 = (size_t)4 * XG * YG * ZG, n2 = (size_t)XG * YG;
    for (size_t e = 0; e < n4; ++e) {
        double a = 0.98 * grid4d[e], b = 0.02 * clim4d[e];
        forecast_4d[e] = a + b;
    }
    for (size_t e = 0; e < n2; ++e) {
        double a = 0.98 * grid2d[e], b = 0.02 * clim2d[e];
        forecast_2d[e] = a + b;
    }
}

/* rolling_average_over_a_period_2d (src/mod_utilities.f90:1773-1815), in place on grid(i,t) = grid[ld*t + i], i < nrows:
 *   t - period < 1 (1-based):  sum(copy(i,1:t)) / t
 *   else:                      sum(copy(i,t-period:t)) / period   (period+1 values over period, as written), kept only
 *                              when |sum| > 1e-7 (keep_small; the 3-D variant :1731-1771 has no such test)
 * windows summed first to last, like the Fortran intrinsic without reassociation. */
void orc_rolling_average_2d(double *grid, int ld, int nrows, int t_len, int period, int keep_small)
{
    double *copy = (double *)malloc(sizeof(double) * (size_t)nrows * (size_t)t_len);
    if (!copy) return;
    for (int t = 0; t < t_len; ++t)
        for (int i = 0; i < nrows; ++i) copy[(size_t)nrows * t + i] = grid[(size_t)ld * t + i];
    for (int i = 0; i < nrows; ++i) {
        for (int t1 = 1; t1 <= t_len; ++t1) {   /* 1-based time index, as in the reference */
            const int head = (t1 - period < 1);
            const int lo = head ? 1 : t1 - period;
            double s = 0.0;
            for (int k = lo; k <= t1; ++k) s = s + copy[(size_t)nrows * (k - 1) + i];
            double out;
            if (head) out = s / (double)t1;
            else if (keep_small && !(fabs(s) > 0.0000001)) out = copy[(size_t)nrows * (t1 - 1) + i];
            else out = s / (double)period;
            grid[(size_t)ld * (t1 - 1) + i] = out;
        }
    }
    free(copy);
}

/* feedback / local_model construction: :581-604 (root), :606-739 (exchange, same tilers), :749-775.
 * tisr_grid is the global 96x48 TISR field (physical units) for this date; the reference keeps the
 * region's slice of the year table pre-standardised with (x-mean)/std (src/mod_reservoir.f90:905-907)
 * and copies the hour's slice (src/mpires.f90:1706) -- same arithmetic per element.
 * sst_mean/sst_std: per-region scalars of the ocean reservoir's grid (grid_special sst_mean_std_idx). */
void orc_step_scatter(orc_region **regs, int nreg, int precip_bool, int ocean_model, int ml_only,
                      const double *wholegrid4d, const double *wholegrid2d, const double *wholegrid_precip,
                      const double *wholegrid_sst, const double *forecast_4d, const double *forecast_2d,
                      const double *tisr_grid, const double *sst_mean, const double *sst_std, int nthreads)
{
    (void)nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (int i = 0; i < nreg; ++i) {
        orc_region *r = regs[i];
        const orc_grid *g = &r->g;
        const orc_dims *d = &r->d;
        const int R = g->number_of_regions, ixy = g->inputxchunk * g->inputychunk;
        orc_tile_4d_and_logp_to_local_state_input(R, g->region, g->overlap, g->num_vert_levels, g->level_index,
                                                  g->vert_overlap, precip_bool, wholegrid4d, wholegrid2d,
                                                  wholegrid_precip, r->feedback);
        if (!ml_only) {
            orc_tile_4d_and_logp_full_grid_to_local_res_vec(R, g->region, g->num_vert_levels, g->level_index,
                                                            forecast_4d, forecast_2d, r->local_model);
            orc_standardize_state_vec_res(g, d, r->mean, r->std, r->local_model);
        }
        if (d->tisr_input_bool) {
            double *t = r->feedback + (g->tisr_start - 1);
            orc_tileoverlapgrid2d(tisr_grid, R, g->region, g->overlap, t);
            const int ti = g->tisr_mean_std_idx - 1;
            for (int e = 0; e < ixy; ++e) {
                double v = t[e] - r->mean[ti];
                t[e] = v / r->std[ti];
            }
        }
        if (ocean_model && d->sst_bool_input) {
            double *s = r->feedback + (g->sst_start - 1);
            orc_tileoverlapgrid2d(wholegrid_sst, R, g->region, g->overlap, s);
            for (int e = 0; e < ixy; ++e) { /* standardize_data_given_pars1d :599,725 */
                double v = s[e] - sst_mean[i];
                s[e] = v / sst_std[i];
            }
        }
        orc_standardize_state_vec_input(g, d, r->mean, r->std, r->feedback); /* :765-769 */
        if (d->precip_bool) { /* :771-773 */
            double *p = r->feedback + (g->precip_start - 1);
            const int pi = g->precip_mean_std_idx - 1;
            for (int e = 0; e < ixy; ++e) {
                double v = p[e] - r->mean[pi];
                p[e] = v / r->std[pi];
            }
        }
    }
}

/* ------------------------------------------------------------------ */
/* training                                                            */
/* ------------------------------------------------------------------ */

/* src/mod_reservoir.f90:1561-1592 initialize_chunk_training (batch_size decided by the caller with
 * orc_find_closest_divisor exactly as :1570-1575) */
int orc_train_init(orc_region *r, int batch_size)
{
    const int n = r->d.n, P = r->d.chunk_size_prediction, S = r->d.chunk_size_speedy;
    const size_t N = (size_t)n + S;
    orc_train_free(r);
    r->batch_size = batch_size;
    r->sxt = (double *)calloc((size_t)P * N, sizeof(double));
    r->sxs = (double *)calloc(N * N, sizeof(double));
    r->states = (double *)calloc((size_t)n * batch_size, sizeof(double));
    r->augmented_states = (double *)calloc(N * batch_size, sizeof(double));
    r->saved_state = (double *)calloc((size_t)n, sizeof(double));
    return (r->sxt && r->sxs && r->states && r->augmented_states && r->saved_state) ? 0 : -1;
}

/* src/mod_reservoir.f90:1645-1701 chunking_matmul (hybrid) / :1594-1642 chunking_matmul_ml.
 * trainingdata / imperfect are the phase's (already strided) series; 1-based batch_number. */
static void chunking_matmul(orc_region *r, int batch_number, const double *trainingdata, int ld_t,
                            const double *imperfect, int ld_i, int discard_cols)
{
    const int n = r->d.n, P = r->d.chunk_size_prediction, S = r->d.chunk_size_speedy;
    const int N = n + S, m = r->batch_size;
    const int col0 = discard_cols + (batch_number - 1) * m; /* 0-based first column */
    double *aug = r->augmented_states;
    for (int c = 0; c < m; ++c) {
        if (S > 0 && imperfect)
            for (int i = 0; i < S; ++i) aug[i + (size_t)N * c] = imperfect[i + (size_t)ld_i * (col0 + c)];
        for (int j = 0; j < n; ++j) aug[S + j + (size_t)N * c] = r->states[j + (size_t)n * c];
    }
    double *target = (double *)malloc(sizeof(double) * (size_t)P * m);
    orc_tile_full_input_to_target_data2d(&r->g, &r->d, trainingdata + (size_t)ld_t * col0, ld_t, m, target);
    /* temp = matmul(targetdata, transpose(aug)); sxt += temp */
    double *temp = (double *)calloc((size_t)P * N, sizeof(double));
    for (int c = 0; c < m; ++c)
        for (int j = 0; j < N; ++j) {
            const double a = aug[j + (size_t)N * c];
            for (int p = 0; p < P; ++p) temp[p + (size_t)P * j] += target[p + (size_t)P * c] * a;
        }
    for (size_t e = 0; e < (size_t)P * N; ++e) r->sxt[e] += temp[e];
    free(temp);
    free(target);
    /* DGEMM('N','N',n,n,m,1,aug,n,transpose(aug),m,0,temp,n); sxs += temp -- full square */
    double *t2 = (double *)calloc((size_t)N * N, sizeof(double));
#pragma omp parallel for schedule(static)
    for (int j = 0; j < N; ++j)
        for (int c = 0; c < m; ++c) {
            const double b = aug[j + (size_t)N * c];
            const double *acol = aug + (size_t)N * c;
            double *tcol = t2 + (size_t)N * j;
            for (int i = 0; i < N; ++i) tcol[i] += acol[i] * b;
        }
    for (size_t e = 0; e < (size_t)N * N; ++e) r->sxs[e] += t2[e];
    free(t2);
}

/* noise is applied by the caller (pre-noised inputs) */
static void train_phase(orc_region *r, const double *trainingdata, int ld_t, const double *imperfect, int ld_i,
                        int ncols, int discard_cols, int hybrid)
{
    const int n = r->d.n, bs = r->batch_size;
    const double leak = r->d.leakage;
    double *x = (double *)calloc((size_t)n, sizeof(double));
    double *y = (double *)calloc((size_t)n, sizeof(double));
    double *temp = (double *)calloc((size_t)n, sizeof(double));
    double *src = (double *)calloc((size_t)n, sizeof(double));
    /* discard loop :1093-1106 */
    for (int i = 1; i <= discard_cols; ++i) state_update(r, trainingdata + (size_t)ld_t * (i - 1), x, y, temp);
    for (int j = 0; j < n; ++j) r->states[j] = x[j]; /* states(:,1) = x */
    int batch_number = 0;
    const int training_length = ncols - discard_cols;
#define STATES_COL(c) (r->states + (size_t)n * ((c)-1))
    for (int i = 1; i <= training_length - 1; ++i) {
        const double *u = trainingdata + (size_t)ld_t * (discard_cols + i - 1);
        if ((i + 1) % bs == 0) {
            batch_number++;
            memcpy(src, STATES_COL(i % bs), sizeof(double) * (size_t)n);
            orc_coo_mv(n, r->d.k, r->rows, r->cols, r->vals, src, y);
            dense_win_gemv(r, u, temp);
            for (int j = 0; j < n; ++j) {
                double x_ = tanh(y[j] + temp[j]);
                double a = (1.0 - leak) * x[j], b = leak * x_;
                x[j] = a + b;
            }
            memcpy(STATES_COL(bs), x, sizeof(double) * (size_t)n);
            memcpy(r->saved_state, STATES_COL(bs), sizeof(double) * (size_t)n);
            for (int c = 0; c < bs; ++c) /* states(2:n:2,:) squared in place :1135 */
                for (int j = 1; j < n; j += 2) {
                    double v = r->states[j + (size_t)n * c];
                    r->states[j + (size_t)n * c] = v * v;
                }
            chunking_matmul(r, batch_number, trainingdata, ld_t, hybrid ? imperfect : NULL, ld_i, discard_cols);
        } else if (i % bs == 0) {
            /* hybrid restarts from saved_state (:1142); ML-only from the squared states(:,batch_size) (:1034) */
            if (hybrid) memcpy(src, r->saved_state, sizeof(double) * (size_t)n);
            else memcpy(src, STATES_COL(bs), sizeof(double) * (size_t)n);
            orc_coo_mv(n, r->d.k, r->rows, r->cols, r->vals, src, y);
            dense_win_gemv(r, u, temp);
            for (int j = 0; j < n; ++j) {
                double x_ = tanh(y[j] + temp[j]);
                double a = (1.0 - leak) * x[j], b = leak * x_;
                x[j] = a + b;
            }
            memcpy(STATES_COL(1), x, sizeof(double) * (size_t)n);
        } else {
            memcpy(src, STATES_COL(i % bs), sizeof(double) * (size_t)n);
            orc_coo_mv(n, r->d.k, r->rows, r->cols, r->vals, src, y);
            dense_win_gemv(r, u, temp);
            for (int j = 0; j < n; ++j) {
                double x_ = tanh(y[j] + temp[j]);
                double a = (1.0 - leak) * x[j], b = leak * x_;
                x[j] = a + b;
            }
            memcpy(STATES_COL((i + 1) % bs), x, sizeof(double) * (size_t)n);
        }
    }
#undef STATES_COL
    free(x); free(y); free(temp); free(src);
}

/* src/mod_reservoir.f90:1067-1175 reservoir_layer_chunking_hybrid */
void orc_train_phase_hybrid(orc_region *r, const double *trainingdata, int ld_t, const double *imperfect,
                            int ld_i, int ncols, int discard_cols)
{
    train_phase(r, trainingdata, ld_t, imperfect, ld_i, ncols, discard_cols, 1);
}

/* src/mod_reservoir.f90:963-1065 reservoir_layer_chunking_ml */
void orc_train_phase_ml(orc_region *r, const double *trainingdata, int ld_t, int ncols, int discard_cols)
{
    train_phase(r, trainingdata, ld_t, NULL, 0, ncols, discard_cols, 0);
}

static int ridge_solve(orc_region *r)
{
    const int n = r->d.n, P = r->d.chunk_size_prediction, S = r->d.chunk_size_speedy;
    const int N = n + S;
    /* a_trans = transpose(sxs); b_trans = transpose(sxt); mldivide; wout = transpose(b_trans) */
    double *a_trans = (double *)malloc(sizeof(double) * (size_t)N * N);
    double *b_trans = (double *)malloc(sizeof(double) * (size_t)N * P);
    for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) a_trans[i + (size_t)N * j] = r->sxs[j + (size_t)N * i];
    for (int p = 0; p < P; ++p)
        for (int i = 0; i < N; ++i) b_trans[i + (size_t)N * p] = r->sxt[p + (size_t)P * i];
    int info = orc_mldivide(a_trans, N, N, b_trans, N, P);
    for (int p = 0; p < P; ++p)
        for (int i = 0; i < N; ++i) r->wout[p + (size_t)P * i] = b_trans[i + (size_t)N * p];
    free(a_trans);
    free(b_trans);
    return info;
}

/* src/mod_reservoir.f90:1235-1334 fit_chunk_hybrid.  prior(i,i) = prior_val*beta_model**2 is added to
 * b_trans^T, i.e. to sxt(i,i) for i <= chunk_size_speedy (:1261-1270,1308-1310). */
int orc_fit_chunk_hybrid(orc_region *r, double beta_res, double beta_model, int using_prior, double prior_val)
{
    const int n = r->d.n, P = r->d.chunk_size_prediction, S = r->d.chunk_size_speedy;
    const int N = n + S;
    for (int i = 0; i < N; ++i) {
        double add;
        if (using_prior) add = (i < S) ? pow(beta_model, 2.0) : pow(beta_res, 2.0);
        else add = (i < S) ? beta_model : beta_res;
        r->sxs[i + (size_t)N * i] += add;
    }
    if (using_prior)
        for (int i = 0; i < S && i < P; ++i) r->sxt[i + (size_t)P * i] += prior_val * pow(beta_model, 2.0);
    return ridge_solve(r);
}

/* src/mod_reservoir.f90:1177-1233 fit_chunk_ml: only the first n diagonals get + beta_res */
int orc_fit_chunk_ml(orc_region *r, double beta_res)
{
    const int n = r->d.n, S = r->d.chunk_size_speedy;
    const int N = n + S;
    for (int i = 0; i < n; ++i) r->sxs[i + (size_t)N * i] += beta_res;
    return ridge_solve(r);
}
