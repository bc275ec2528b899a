// replay_main.cpp -- a compiled host program on the C++ host layer (speedy-ml_b200/host/speedyml_host.hpp).
//
// Replays the prediction part of src/parallelmain.f90 (:143-273) on a case file: load every region's trained
// reservoir (trained_reservoir_prediction -> mklsparse), synchronize, then the hybrid loop
//     do t: predict(all regions); sendrecievegrid(res, t, slab_model)      [run_model = deterministic stand-in]
// and writes the grids of every step.  tests/test_cpp_host_gpu.py builds the case, runs this binary on the GPU and
// checks the grids against the CPU oracle and against the Python host path.
//     replay_main <case file> <output file> [--overlap]
#include "../host/speedyml_host.hpp"

#include <chrono>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>

using namespace speedyml;

namespace {
template <typename T>
void rd(std::ifstream &f, T *p, size_t n)
{
    f.read(reinterpret_cast<char *>(p), sizeof(T) * n);
    if (!f) throw std::runtime_error("case file truncated");
}
template <typename T>
std::vector<T> rdv(std::ifstream &f, size_t n)
{
    std::vector<T> v(n);
    if (n) rd(f, v.data(), n);
    return v;
}
}  // namespace

int main(int argc, char **argv)
{
    if (argc < 3) {
        std::cerr << "usage: replay_main <case> <out> [--overlap]\n";
        return 2;
    }
    const bool overlap = argc > 3 && std::strcmp(argv[3], "--overlap") == 0;
    try {
        std::ifstream f(argv[1], std::ios::binary);
        if (!f) throw std::runtime_error("cannot open case file");
        char magic[8];
        rd(f, magic, 8);
        if (std::memcmp(magic, "SMLCASE1", 8)) throw std::runtime_error("not a case file");
        int32_t hdr[9];
        rd(f, hdr, 9);
        model_parameters_type mp;
        mp.number_of_regions = hdr[0];
        mp.overlap = hdr[1];
        mp.precip_bool = hdr[2];
        mp.slab_ocean_model_bool = hdr[3];
        mp.ml_only = hdr[4];
        mp.sst_prescribed = hdr[5];
        const int nreg = hdr[6], nsteps = hdr[7], sync_len = hdr[8];
        Engine eng(mp);
        if ((int)mp.region_indices.size() != nreg) throw std::runtime_error("case does not hold this rank's regions");

        std::vector<reservoir_type> res(nreg);
        std::vector<grid_type> grid(nreg);
        std::vector<std::vector<dp>> sync_in(nreg);
        for (int i = 0; i < nreg; ++i) {
            reservoir_type &r = res[i];
            int32_t d[8];
            rd(f, d, 8);
            r.assigned_region = d[0]; r.n = d[1]; r.k = d[2]; r.reservoir_numinputs = d[3];
            r.chunk_size_prediction = d[4]; r.chunk_size_speedy = d[5];
            const int L = d[6];
            r.sst_bool_input = d[7];
            rd(f, &r.leakage, 1);
            r.rows = rdv<int32_t>(f, r.k);
            r.cols = rdv<int32_t>(f, r.k);
            r.vals = rdv<dp>(f, r.k);
            r.win_compact = rdv<dp>(f, r.n);
            r.win_col = rdv<int32_t>(f, r.n);
            r.wout = rdv<dp>(f, (size_t)r.chunk_size_prediction * (r.n + r.chunk_size_speedy));
            grid[i].mean = rdv<dp>(f, L);
            grid[i].std = rdv<dp>(f, L);
            grid[i].sst_mean_std_idx = mp.slab_ocean_model_bool ? L : 0;  // the SST slot is the last one
            r.saved_state = rdv<dp>(f, r.n);
            r.feedback = rdv<dp>(f, r.reservoir_numinputs);
            r.local_model = rdv<dp>(f, r.chunk_size_speedy);
            sync_in[i] = rdv<dp>(f, (size_t)r.reservoir_numinputs * sync_len);
            eng.mklsparse(r, grid[i]);   // trained_reservoir_prediction -> mklsparse
        }
        const size_t n4 = 4 * SML_XGRID * SML_YGRID * SML_ZGRID, n2 = SML_XGRID * SML_YGRID;
        std::vector<dp> clim4d = rdv<dp>(f, n4), clim2d = rdv<dp>(f, n2), tisr = rdv<dp>(f, n2);
        mp.base_sst_grid = rdv<dp>(f, n2);
        mp.sea_mask = rdv<dp>(f, n2);
        eng.finalize(mp);
        if (mp.sst_prescribed) eng.set_sst_prescribed(mp.base_sst_grid.data());

        // start_prediction (src/mod_reservoir.f90:940-961): synchronize on the recent data, current_state = saved_state
        for (int i = 0; i < nreg; ++i) {
            reservoir_type &r = res[i];
            if (sync_len > 0) eng.synchronize(r, sync_in[i].data(), r.reservoir_numinputs, r.saved_state, sync_len);
            r.current_state = r.saved_state;
            eng.start_prediction(r);
        }

        // run_model stand-in (SPEEDY stays on the host): forecast = 0.98*grid + 0.02*climatology, q floor 1e-6
        Engine::run_model_fn run_model = [&](int, std::vector<dp> &g4, std::vector<dp> &g2, std::vector<dp> &, std::vector<dp> &f4,
                                             std::vector<dp> &f2) {
            for (size_t e = 0; e < n4; ++e) {
                const dp a = 0.98 * g4[e], b = 0.02 * clim4d[e];
                f4[e] = a + b;
            }
            for (size_t p = 0; p < n4 / 4; ++p)
                if (f4[3 + 4 * p] < 0.000001) f4[3 + 4 * p] = 0.000001;
            for (size_t e = 0; e < n2; ++e) {
                const dp a = 0.98 * g2[e], b = 0.02 * clim2d[e];
                f2[e] = a + b;
            }
        };

        std::ofstream out(argv[2], std::ios::binary);
        Engine::grids G;
        if (overlap) eng.set_overlap(true);
        const auto t0 = std::chrono::steady_clock::now();
        for (int t = 1; t <= nsteps; ++t) {
            // region loop of src/parallelmain.f90:226-251: the first predict of the step runs every local region
            for (int i = 0; i < nreg; ++i) eng.predict(res[i]);
            if (overlap) eng.set_tisr(tisr.data());
            eng.sendrecievegrid(mp, t, run_model, tisr.data(), G);
            out.write(reinterpret_cast<const char *>(G.wholegrid4d.data()), sizeof(dp) * n4);
            out.write(reinterpret_cast<const char *>(G.wholegrid2d.data()), sizeof(dp) * n2);
            out.write(reinterpret_cast<const char *>(G.wholegrid_precip.data()), sizeof(dp) * n2);
            out.write(reinterpret_cast<const char *>(G.wholegrid_sst.data()), sizeof(dp) * n2);
        }
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        // final outvec and feedback of every region
        for (int i = 0; i < nreg; ++i) {
            out.write(reinterpret_cast<const char *>(res[i].outvec.data()), sizeof(dp) * res[i].outvec.size());
            eng.feedback_get(res[i]);
            out.write(reinterpret_cast<const char *>(res[i].feedback.data()), sizeof(dp) * res[i].feedback.size());
        }
        std::printf("replay ok: %d regions, %d steps, %.3f ms per step, %lld kernel launches%s\n", nreg, nsteps,
                    1e3 * secs / (nsteps > 0 ? nsteps : 1), eng.kernel_launch_count(), overlap ? " (overlapped)" : "");
        std::printf("program finished correctly\n");
    } catch (const std::exception &e) {
        std::fprintf(stderr, "replay_main: %s\n", e.what());
        return 1;
    }
    return 0;
}
