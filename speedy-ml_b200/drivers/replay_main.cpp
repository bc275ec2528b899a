// replay_main.cpp -- a compiled host program on the C++ host layer (speedy-ml_b200/host/speedyml_host.hpp).
//
// Replays the prediction part of src/parallelmain.f90 (:143-273) on a case file: load every region's trained
// reservoir (trained_reservoir_prediction -> mklsparse), synchronize, then the hybrid loop
//     do t: predict(all regions); sendrecievegrid(res, t, slab_model)      [run_model = deterministic stand-in]
// and writes the grids of every step.  tests/test_cpp_host_gpu.py builds the case, runs this binary on the GPU and
// checks the grids against the CPU oracle and against the Python host path.
//     replay_main <case file> <output file> [--overlap] [--batched-sync]
//                 [--rank R --world W --shm NAME [--device D]]
// Multi-rank (one process per GPU, as the reference runs one MPI rank per region block): the case file then holds
// EVERY region of the model and each rank keeps its own (processor_decomposition); the ranks connect through
// sml_comm_bootstrap over a POSIX shared-memory all-gather (ShmAllgather stands in for MPI_Allgather) and from then on
// only call predict / sendrecievegrid -- rank 0 with the host model, the others without ever blocking on it.  Rank 0
// writes every step's grids; every rank appends the grids IT assembled in the last step (all ranks rebuild the whole
// grid) and its regions' final outvec and feedback.
#include "../host/speedyml_host.hpp"

#include <chrono>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <unordered_map>

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
        std::cerr << "usage: replay_main <case> <out> [--overlap] [--batched-sync] [--rank R --world W --shm NAME [--device D]]\n";
        return 2;
    }
    bool overlap = false, batched_sync = false;
    int rank = 0, world = 1, device = -1;
    std::string shm_name;
    for (int a = 3; a < argc; ++a) {
        const std::string arg = argv[a];
        if (arg == "--overlap") overlap = true;
        else if (arg == "--batched-sync") batched_sync = true;
        else if (arg == "--rank" && a + 1 < argc) rank = std::atoi(argv[++a]);
        else if (arg == "--world" && a + 1 < argc) world = std::atoi(argv[++a]);
        else if (arg == "--device" && a + 1 < argc) device = std::atoi(argv[++a]);
        else if (arg == "--shm" && a + 1 < argc) shm_name = argv[++a];
        else {
            std::cerr << "replay_main: unknown argument " << arg << "\n";
            return 2;
        }
    }
    if (device < 0) device = rank;
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
        mp.irank = rank;
        mp.numprocs = world;
        const int nrec = hdr[6], nsteps = hdr[7], sync_len = hdr[8];
        Engine eng(mp, device);
        const int nreg = (int)mp.region_indices.size();
        if (world == 1 && nrec != nreg) throw std::runtime_error("case does not hold this rank's regions");
        std::unordered_map<int, int> mine;
        for (int i = 0; i < nreg; ++i) mine[mp.region_indices[i]] = i;

        std::vector<reservoir_type> res(nreg);
        std::vector<grid_type> grid(nreg);
        std::vector<std::vector<dp>> sync_in(nreg);
        int loaded = 0;
        for (int rec = 0; rec < nrec; ++rec) {
            reservoir_type tmp;
            grid_type gtmp;
            int32_t d[8];
            rd(f, d, 8);
            auto it = mine.find(d[0]);
            const bool keep = it != mine.end();
            reservoir_type &r = keep ? res[it->second] : tmp;
            grid_type &g = keep ? grid[it->second] : gtmp;
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
            g.mean = rdv<dp>(f, L);
            g.std = rdv<dp>(f, L);
            g.sst_mean_std_idx = mp.slab_ocean_model_bool ? L : 0;  // the SST slot is the last one
            r.saved_state = rdv<dp>(f, r.n);
            r.feedback = rdv<dp>(f, r.reservoir_numinputs);
            r.local_model = rdv<dp>(f, r.chunk_size_speedy);
            std::vector<dp> si = rdv<dp>(f, (size_t)r.reservoir_numinputs * sync_len);
            if (!keep) continue;
            sync_in[it->second] = std::move(si);
            eng.mklsparse(r, g);   // trained_reservoir_prediction -> mklsparse
            ++loaded;
        }
        if (loaded != nreg) throw std::runtime_error("case file does not hold every region of this rank");
        const size_t n4 = 4 * SML_XGRID * SML_YGRID * SML_ZGRID, n2 = SML_XGRID * SML_YGRID;
        std::vector<dp> clim4d = rdv<dp>(f, n4), clim2d = rdv<dp>(f, n2), tisr = rdv<dp>(f, n2);
        mp.base_sst_grid = rdv<dp>(f, n2);
        mp.sea_mask = rdv<dp>(f, n2);
        eng.finalize(mp);
        if (mp.sst_prescribed) eng.set_sst_prescribed(mp.base_sst_grid.data());

        // the ranks connect: ONE host primitive (an all-gather of a few bytes), everything else is the engine's
        std::unique_ptr<ShmAllgather> shm;
        if (world > 1) {
            if (shm_name.empty()) throw std::runtime_error("--world > 1 needs --shm NAME");
            shm.reset(new ShmAllgather(shm_name, rank, world));
            eng.comm_bootstrap(&ShmAllgather::callback, shm.get());
        }

        // start_prediction (src/mod_reservoir.f90:940-961): synchronize on the recent data, current_state = saved_state
        if (batched_sync && sync_len > 0) {
            std::vector<reservoir_type *> local;
            std::vector<const dp *> inputs;
            for (int i = 0; i < nreg; ++i) {
                local.push_back(&res[i]);
                inputs.push_back(sync_in[i].data());
            }
            eng.synchronize_all(local, inputs, sync_len);   // one launch per time step for all regions
        }
        for (int i = 0; i < nreg; ++i) {
            reservoir_type &r = res[i];
            if (!batched_sync && sync_len > 0) eng.synchronize(r, sync_in[i].data(), r.reservoir_numinputs, r.saved_state, sync_len);
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
        auto write_grids = [&]() {
            out.write(reinterpret_cast<const char *>(G.wholegrid4d.data()), sizeof(dp) * n4);
            out.write(reinterpret_cast<const char *>(G.wholegrid2d.data()), sizeof(dp) * n2);
            out.write(reinterpret_cast<const char *>(G.wholegrid_precip.data()), sizeof(dp) * n2);
            out.write(reinterpret_cast<const char *>(G.wholegrid_sst.data()), sizeof(dp) * n2);
        };
        if (overlap) eng.set_overlap(true);
        if (shm) shm->barrier();
        const auto t0 = std::chrono::steady_clock::now();
        for (int t = 1; t <= nsteps; ++t) {
            // region loop of src/parallelmain.f90:226-251: the first predict of the step runs every local region
            for (int i = 0; i < nreg; ++i) eng.predict(res[i]);
            if (overlap) eng.set_tisr(tisr.data());
            eng.sendrecievegrid(mp, t, run_model, tisr.data(), G);
            if (rank == 0) write_grids();
        }
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (world > 1) {
            // every rank rebuilt the whole grid of the last step: all copies must be identical
            eng.grids_get(nsteps, G);
            write_grids();
        }
        // final outvec and feedback of every local region
        for (int i = 0; i < nreg; ++i) {
            out.write(reinterpret_cast<const char *>(res[i].outvec.data()), sizeof(dp) * res[i].outvec.size());
            eng.feedback_get(res[i]);
            out.write(reinterpret_cast<const char *>(res[i].feedback.data()), sizeof(dp) * res[i].feedback.size());
        }
        if (shm) shm->barrier();   // nobody tears its exchange block down while a peer may still push into it
        std::printf("replay ok: rank %d of %d, %d regions, %d steps, %.3f ms per step, %lld kernel launches%s\n", rank, world, nreg,
                    nsteps, 1e3 * secs / (nsteps > 0 ? nsteps : 1), eng.kernel_launch_count(), overlap ? " (overlapped)" : "");
        std::printf("program finished correctly\n");
    } catch (const std::exception &e) {
        std::fprintf(stderr, "replay_main: %s\n", e.what());
        return 1;
    }
    return 0;
}
