// speedyml_host.hpp -- C++ host layer over the C ABI (include/speedyml_engine.h).
//
// The reference is compiled code (Fortran 90): its hot path is reached through module procedures on derived types.
// This header mirrors that call surface in C++ -- the same procedure names, the same argument meaning, the same
// error behaviour -- so that a compiled host can drive the engine without Python:
//     mklsparse            src/mod_linalg.f90:10          synchronize      src/mod_reservoir.f90:1354
//     predict / predict_ml src/mod_reservoir.f90:1418,1491 predict_slab_ml  src/mod_slab_ocean_reservoir.f90:1318
//     sendrecievegrid      src/mpires.f90:218             mldivide         src/mod_linalg.f90:109
//     train_reservoir's inner sequence (initialize_chunk_training / reservoir_layer_chunking_* / fit_chunk_*)
//     gen_res              src/mod_reservoir.f90:182
// The derived types below carry the fields of reservoir_type / grid_type / model_parameters_type
// (src/mod_utilities.f90:32-509) that the path reads or writes; arrays are column-major like the Fortran ones.
// The Fortran twin of this layer is speedy-ml_b200/fortran/speedyml_gpu.f90 (uncompiled here: no Fortran compiler);
// speedy-ml_b200/drivers/replay_main.cpp is a host program built on this header and run by the GPU tests.
#pragma once
#include "../../include/speedyml_engine.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace speedyml {

using dp = double;  // real(kind=dp)

struct model_parameters_type {
    int number_of_regions = 1152, overlap = 1;
    bool precip_bool = true, slab_ocean_model_bool = true, ml_only = false, ml_only_ocean = true;
    bool outvec_component_contribs = false;
    int irank = 0, numprocs = 1, timestep = 6, timestep_slab = 168;
    bool sst_prescribed = false;               // engine extension: SST field supplied by the host every step
    std::vector<int> region_indices;           // filled by the engine (processor_decomposition)
    std::vector<dp> base_sst_grid, sea_mask;   // (xgrid, ygrid)
    bool run_speedy = true;
};

struct grid_type {
    std::vector<dp> mean, std;                 // grid%mean / grid%std
    int sst_mean_std_idx = 0;                  // 1-based, 0 if absent
};

struct reservoir_type {
    int assigned_region = 0;
    int n = 0, k = 0, reservoir_numinputs = 0, chunk_size_prediction = 0, chunk_size_speedy = 0;
    bool sst_bool_input = true, sst_bool_prediction = true;
    dp leakage = 1.0, beta_res = 0.001, beta_model = 1.0, prior_val = 0.0, radius = 0.7;
    std::vector<int> rows, cols;               // 1-based COO
    std::vector<dp> vals;
    std::vector<dp> win;                       // dense (n, reservoir_numinputs), or empty with the compact form
    std::vector<dp> win_compact;               // the single non-zero of each row
    std::vector<int> win_col;                  // its 0-based column
    std::vector<dp> wout;                      // (chunk_size_prediction, n + chunk_size_speedy)
    std::vector<dp> feedback, local_model, outvec, current_state, saved_state, v_p, v_ml;
};

// The one primitive sml_comm_bootstrap asks of the host, for ranks that are processes of ONE node and have no MPI at
// hand (the reference would pass MPI_Allgather on mpi_res%mpi_world): an all-gather through a POSIX shared-memory
// segment.  Rank 0 creates the segment, every rank writes its block and waits until all `world` blocks of the round
// are in.  Rounds alternate between two halves of the segment so that a fast rank cannot overwrite a slow one's read.
class ShmAllgather {
public:
    static constexpr int MAX_BYTES = 512;
    ShmAllgather(const std::string &name, int rank, int world) : name_("/" + name), rank_(rank), world_(world)
    {
        const size_t bytes = sizeof(Header) + 2 * (size_t)world * MAX_BYTES;
        int fd = -1;
        if (rank == 0) {
            shm_unlink(name_.c_str());
            fd = shm_open(name_.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
            if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) throw std::runtime_error("ShmAllgather: cannot create " + name_);
        } else {
            for (int tries = 0; tries < 60000; ++tries) {   // up to ~60 s for rank 0 to get there
                fd = shm_open(name_.c_str(), O_RDWR, 0600);
                struct stat st;
                if (fd >= 0 && fstat(fd, &st) == 0 && (size_t)st.st_size >= bytes) break;
                if (fd >= 0) { close(fd); fd = -1; }
                usleep(1000);
            }
            if (fd < 0) throw std::runtime_error("ShmAllgather: rank 0 never created " + name_);
        }
        void *m = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (m == MAP_FAILED) throw std::runtime_error("ShmAllgather: mmap failed");
        base_ = static_cast<unsigned char *>(m);
        bytes_ = bytes;
    }
    ~ShmAllgather()
    {
        if (base_) munmap(base_, bytes_);
        if (rank_ == 0) shm_unlink(name_.c_str());
    }
    ShmAllgather(const ShmAllgather &) = delete;
    ShmAllgather &operator=(const ShmAllgather &) = delete;

    int allgather(const void *send, void *recv, int nbytes)
    {
        if (nbytes > MAX_BYTES) return 1;
        Header *hd = reinterpret_cast<Header *>(base_);
        unsigned char *half = base_ + sizeof(Header) + (size_t)(round_ & 1) * world_ * MAX_BYTES;
        std::memcpy(half + (size_t)rank_ * MAX_BYTES, send, (size_t)nbytes);
        hd->arrived.fetch_add(1, std::memory_order_acq_rel);
        const unsigned want = (unsigned)world_ * (unsigned)(round_ + 1);
        for (long spins = 0; hd->arrived.load(std::memory_order_acquire) < want; ++spins) {
            if (spins > 120000) return 1;   // ~2 min: a rank died
            usleep(1000);
        }
        for (int k = 0; k < world_; ++k)
            std::memcpy(static_cast<unsigned char *>(recv) + (size_t)k * nbytes, half + (size_t)k * MAX_BYTES, (size_t)nbytes);
        ++round_;
        return 0;
    }
    // signature of sml_allgather_fn; ctx is the ShmAllgather
    static int callback(void *ctx, const void *send, void *recv, int nbytes)
    {
        return static_cast<ShmAllgather *>(ctx)->allgather(send, recv, nbytes);
    }
    void barrier()
    {
        int32_t one = 1;
        std::vector<int32_t> all((size_t)world_);
        if (allgather(&one, all.data(), (int)sizeof(one))) throw std::runtime_error("ShmAllgather: barrier timed out");
    }

private:
    struct Header {
        std::atomic<unsigned> arrived;
        unsigned pad[15];
    };
    std::string name_;
    int rank_, world_;
    unsigned char *base_ = nullptr;
    size_t bytes_ = 0;
    int round_ = 0;
};

class Engine {
public:
    static constexpr int ATMO = SML_ATMO, OCEAN = SML_OCEAN;

    explicit Engine(model_parameters_type &mp, int device = 0)
    {
        sml_params p{};
        p.number_of_regions = mp.number_of_regions;
        p.overlap = mp.overlap;
        p.precip_bool = mp.precip_bool;
        p.slab_ocean_model_bool = mp.slab_ocean_model_bool;
        p.ml_only = mp.ml_only;
        p.irank = mp.irank;
        p.numprocs = mp.numprocs;
        p.device = device;
        p.timestep = mp.timestep;
        p.timestep_slab = mp.timestep_slab;
        p.sst_prescribed = mp.sst_prescribed;
        if (sml_create(&h_, &p)) throw std::runtime_error(std::string("sml_create: ") + sml_last_error(nullptr));
        mp.region_indices.resize(sml_num_local_regions(h_));
        sml_local_region_ids(h_, mp.region_indices.data());
        for (size_t i = 0; i < mp.region_indices.size(); ++i) local_of_[mp.region_indices[i]] = (int)i;
        contribs_ = mp.outvec_component_contribs;
        irank_ = mp.irank;
        numprocs_ = mp.numprocs;
    }
    ~Engine() { sml_destroy(h_); }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;

    // mklsparse(reservoir): takes the whole reservoir (adjacency, W_in, W_out) and the grid's mean/std, because this
    // is where the reference has all of them in hand (trained_reservoir_prediction, src/mod_reservoir.f90:1852).
    // A failed sparse create makes the reference print and stop (src/mod_linalg.f90:18-22): here it throws.
    void mklsparse(const reservoir_type &r, const grid_type &g, int kind = ATMO)
    {
        sml_region_weights w{};
        w.region = r.assigned_region;
        w.kind = kind;
        w.n = r.n;
        w.k = r.k;
        w.D = r.reservoir_numinputs;
        w.P = r.chunk_size_prediction;
        w.S = r.chunk_size_speedy;
        w.L = (int)g.mean.size();
        w.sst_bool_input = r.sst_bool_input;
        w.leakage = r.leakage;
        w.sst_mean = g.sst_mean_std_idx > 0 ? g.mean[g.sst_mean_std_idx - 1] : 0.0;
        w.sst_std = g.sst_mean_std_idx > 0 ? g.std[g.sst_mean_std_idx - 1] : 1.0;
        w.rows = r.rows.data();
        w.cols = r.cols.data();
        w.vals = r.vals.data();
        w.win_dense = r.win.empty() ? nullptr : r.win.data();
        w.win_compact = r.win.empty() ? r.win_compact.data() : nullptr;
        w.win_col = r.win.empty() ? r.win_col.data() : nullptr;
        w.wout = r.wout.empty() ? nullptr : r.wout.data();
        w.mean = g.mean.data();
        w.std = g.std.data();
        ck(sml_region_upload(h_, &w), "mklsparse");
    }

    // after the last mklsparse of the rank (end of the load loop, src/parallelmain.f90:160-185)
    void finalize(const model_parameters_type &mp)
    {
        ck(sml_finalize(h_), "sml_finalize");
        if (mp.slab_ocean_model_bool) ck(sml_set_sst_static(h_, mp.base_sst_grid.data(), mp.sea_mask.data()), "sml_set_sst_static");
        if (contribs_) ck(sml_set_contribs(h_, 1), "sml_set_contribs");
    }

    // multi-rank runs (numprocs > 1), once after finalize on every rank: the host hands over its all-gather (the
    // reference's MPI communicator, or ShmAllgather above) and the engine connects the ranks' exchange blocks; from
    // then on sendrecievegrid is complete on every rank without another host collective
    void comm_bootstrap(sml_allgather_fn allgather, void *ctx) { ck(sml_comm_bootstrap(h_, allgather, ctx), "sml_comm_bootstrap"); }

    // gen_res: spectral radius (sparse_eigen) and vals = vals/eig*radius on the device copy and on reservoir%vals
    void gen_res(std::vector<reservoir_type *> &local, int kind = ATMO, int maxit = 500, double tol = 1e-13)
    {
        std::vector<double> eigs(local.size()), factor(local.size(), 1.0);
        int it = 0;
        if (ck(sml_sparse_eigen(h_, kind, maxit, tol, eigs.data(), &it), "sparse_eigen") == 1)
            throw std::runtime_error("sparse_eigen did not converge");
        for (size_t i = 0; i < local.size(); ++i)
            if (eigs[i] > 0.0) factor[i] = local[i]->radius / eigs[i];
        ck(sml_adjacency_scale(h_, kind, factor.data()), "sml_adjacency_scale");
        for (size_t i = 0; i < local.size(); ++i)
            for (double &v : local[i]->vals) v *= factor[i];
    }

    // synchronize(reservoir, input(:,:), x(:), length)
    void synchronize(reservoir_type &r, const dp *input, int ld, std::vector<dp> &x, int length, int kind = ATMO)
    {
        ck(sml_state_set(h_, kind, r.assigned_region, x.data()), "sml_state_set");
        ck(sml_synchronize(h_, kind, r.assigned_region, input, ld, length, nullptr), "synchronize");
        ck(sml_state_get(h_, kind, r.assigned_region, x.data()), "sml_state_get");
    }

    // the same for EVERY local region in one call (the loop over regions of initialize_prediction / start_prediction,
    // src/mod_reservoir.f90:818-824, :951): one batched launch per time step instead of one per region and step.
    // local[i] is the reservoir of mp.region_indices[i]; inputs[i] its (reservoir_numinputs, length) series; the
    // states start from and return to local[i]->saved_state.
    void synchronize_all(std::vector<reservoir_type *> &local, const std::vector<const dp *> &inputs, int length, int kind = ATMO)
    {
        std::vector<dp> flat;
        std::vector<int64_t> offs(local.size(), 0);
        for (size_t i = 0; i < local.size(); ++i) {
            offs[i] = (int64_t)flat.size();
            flat.insert(flat.end(), inputs[i], inputs[i] + (size_t)local[i]->reservoir_numinputs * length);
            ck(sml_state_set(h_, kind, local[i]->assigned_region, local[i]->saved_state.data()), "sml_state_set");
        }
        ck(sml_synchronize(h_, kind, SML_ALL_REGIONS, flat.data(), 0, length, offs.data()), "synchronize");
        for (size_t i = 0; i < local.size(); ++i)
            ck(sml_state_get(h_, kind, local[i]->assigned_region, local[i]->saved_state.data()), "sml_state_get");
    }

    // start_prediction: reservoir%current_state = reservoir%saved_state; feedback / local_model as the host set them
    void start_prediction(reservoir_type &r, int kind = ATMO)
    {
        ck(sml_state_set(h_, kind, r.assigned_region, r.current_state.data()), "sml_state_set");
        ck(sml_feedback_set(h_, kind, r.assigned_region, r.feedback.data()), "sml_feedback_set");
        if (kind == ATMO && !r.local_model.empty())
            ck(sml_local_model_set(h_, kind, r.assigned_region, r.local_model.data()), "sml_local_model_set");
        if (kind == OCEAN) ck(sml_outvec_set(h_, kind, r.assigned_region, r.outvec.data()), "sml_outvec_set");
    }

    // predict(reservoir, model_parameters, grid, x, local_model_in): the first call of a hybrid step launches ONE
    // batched kernel for every local region; later calls of the step only fetch that region's outvec
    void predict(reservoir_type &r)
    {
        if (step_predicted_ != current_step_) {
            ck(sml_predict(h_, ATMO), "predict");
            step_predicted_ = current_step_;
            // ... and ONE read-back of every local outvec (a blocking 1 KB copy per region would cost more than the step)
            outvec_cache_.resize(local_of_.size() * (size_t)r.chunk_size_prediction);
            if (fetch_outvecs_) ck(sml_outvec_get_all(h_, ATMO, outvec_cache_.data()), "sml_outvec_get_all");
        }
        r.outvec.resize(r.chunk_size_prediction);
        if (fetch_outvecs_) {
            const dp *src = outvec_cache_.data() + (size_t)local_of_.at(r.assigned_region) * r.chunk_size_prediction;
            r.outvec.assign(src, src + r.chunk_size_prediction);
        }
        if (contribs_) {
            r.v_p.resize(r.chunk_size_prediction);
            r.v_ml.resize(r.chunk_size_prediction);
            ck(sml_contribs_get(h_, r.assigned_region, r.v_p.data(), r.v_ml.data()), "sml_contribs_get");
        }
    }
    void predict_ml(reservoir_type &r) { predict(r); }
    void predict_all() { ck(sml_predict(h_, ATMO), "predict"); step_predicted_ = current_step_; }
    // reservoir%outvec is only read by the exchange, which lives on the device: a host that does not look at it can
    // switch the per-step read-back off and the predict calls never block
    void set_fetch_outvecs(bool on) { fetch_outvecs_ = on; }

    // predict_slab_ml: the caller keeps the schedule test mod(t*timestep, timestep_slab) == 0 (parallelmain.f90:238)
    void predict_slab_ml(reservoir_type &r)
    {
        if (ocean_step_predicted_ != current_step_) {
            ck(sml_predict(h_, OCEAN), "predict_slab_ml");
            ocean_step_predicted_ = current_step_;
        }
        r.outvec.resize(r.chunk_size_prediction);
        ck(sml_outvec_get(h_, OCEAN, r.assigned_region, r.outvec.data()), "sml_outvec_get");
    }

    // sendrecievegrid(res, timestep, ocean_model): gather + clamps on the device, the host model (run_model,
    // src/mpires.f90:1548-1660) and the output writer stay with the caller, feedback / local_model rebuilt on the
    // device.  run_model(timestep, wholegrid4d, wholegrid2d, wholegrid_sst, forecast_4d, forecast_2d).
    using run_model_fn = std::function<void(int, std::vector<dp> &, std::vector<dp> &, std::vector<dp> &, std::vector<dp> &,
                                            std::vector<dp> &)>;
    struct grids {
        std::vector<dp> wholegrid4d = std::vector<dp>(4 * SML_XGRID * SML_YGRID * SML_ZGRID), wholegrid2d = std::vector<dp>(SML_XGRID * SML_YGRID),
                        wholegrid_precip = std::vector<dp>(SML_XGRID * SML_YGRID), wholegrid_sst = std::vector<dp>(SML_XGRID * SML_YGRID),
                        forecast_4d = std::vector<dp>(4 * SML_XGRID * SML_YGRID * SML_ZGRID), forecast_2d = std::vector<dp>(SML_XGRID * SML_YGRID);
    };
    // Multi-rank (after comm_bootstrap): the root (irank 0) does exactly the above; every other rank only enqueues the
    // grid assembly and the wait for the root's forecast block -- it never blocks on the host model -- and
    // mp.run_speedy follows the root's flag (MPI_Bcast of run_speedy, src/mpires.f90:744) when check_run_speedy is set.
    void sendrecievegrid(model_parameters_type &mp, int timestep, const run_model_fn &run_model, const dp *tisr_grid, grids &G,
                         bool check_run_speedy = false)
    {
        if (numprocs_ > 1 && irank_ != 0) {
            ck(sml_step_exchange_begin(h_, timestep, nullptr, nullptr, nullptr, nullptr), "sendrecievegrid/begin");
            ck(sml_step_exchange_end(h_, timestep, nullptr, nullptr, nullptr), "sendrecievegrid/end");
        } else {
            const int rc = ck(sml_step_exchange_begin(h_, timestep, G.wholegrid4d.data(), G.wholegrid2d.data(),
                                                      G.wholegrid_precip.data(), G.wholegrid_sst.data()), "sendrecievegrid/begin");
            if (rc > 0) mp.run_speedy = false;   // a non-finite grid: SPEEDY must not be run on it
            if (!mp.ml_only && mp.run_speedy)
                run_model(timestep, G.wholegrid4d, G.wholegrid2d, G.wholegrid_sst, G.forecast_4d, G.forecast_2d);
            ck(sml_set_run_speedy(h_, mp.run_speedy ? 1 : 0), "sml_set_run_speedy");
            ck(sml_step_exchange_end(h_, timestep, G.forecast_4d.data(), G.forecast_2d.data(), tisr_grid), "sendrecievegrid/end");
        }
        if (check_run_speedy) {
            int flag = 1;
            ck(sml_run_speedy(h_, &flag), "sml_run_speedy");
            mp.run_speedy = flag != 0;
        }
        current_step_ = timestep + 1;
    }
    // the assembled grids on ANY rank (every rank rebuilds the whole grid): device -> host copy of the current step
    void grids_get(int timestep_just_exchanged, grids &G)
    {
        (void)timestep_just_exchanged;
        ck(sml_grids_get(h_, G.wholegrid4d.data(), G.wholegrid2d.data(), G.wholegrid_precip.data(), G.wholegrid_sst.data()), "sml_grids_get");
    }
    void set_sst_prescribed(const dp *sst) { ck(sml_set_sst_prescribed(h_, sst), "sml_set_sst_prescribed"); }
    void set_overlap(bool on) { ck(sml_set_overlap(h_, on), "sml_set_overlap"); }
    void set_tisr(const dp *tisr) { ck(sml_set_tisr(h_, tisr), "sml_set_tisr"); }

    // train_reservoir's inner sequence for a wave of regions (src/mod_reservoir.f90:287-316)
    void train_begin(const std::vector<int32_t> &regions, int batch_size, int kind = ATMO)
    {
        ck(sml_train_begin(h_, kind, regions.data(), (int)regions.size(), batch_size), "initialize_chunk_training");
    }
    void train_phase(const dp *trainingdata, const int64_t *td_off, const dp *imperfect_model, const int64_t *im_off,
                     int ncols, int discard_cols)
    {
        ck(sml_train_feed(h_, trainingdata, td_off, imperfect_model, im_off, ncols, discard_cols), "reservoir_layer_chunking");
    }
    // fit_chunk_hybrid / fit_chunk_ml: 'something went wrong with dgesv' is print-and-continue (mod_linalg.f90:147-150)
    std::vector<int32_t> fit_chunk(std::vector<reservoir_type *> &wave, bool using_prior, int kind = ATMO)
    {
        std::vector<int32_t> info(wave.size(), 0);
        ck(sml_train_solve(h_, wave[0]->beta_res, wave[0]->beta_model, using_prior, wave[0]->prior_val, info.data()), "fit_chunk");
        for (size_t i = 0; i < wave.size(); ++i) {
            if (info[i] != 0) {
                std::printf("something went wrong with dgesv info = %d\nB is not the solution\n", info[i]);
                continue;
            }
            reservoir_type &r = *wave[i];
            r.wout.resize((size_t)r.chunk_size_prediction * (r.n + r.chunk_size_speedy));
            ck(sml_wout_get(h_, kind, r.assigned_region, r.wout.data()), "sml_wout_get");
        }
        return info;
    }
    void train_end() { ck(sml_train_end(h_), "sml_train_end"); }

    // mldivide(A, B): A X = B, B becomes X when info == 0; returns dgesv's info
    int mldivide(std::vector<dp> &A, int n, std::vector<dp> &B, int nrhs)
    {
        if ((int)(A.size() / (n > 0 ? n : 1)) != n || (int)(B.size() / (nrhs > 0 ? nrhs : 1)) != n) {
            std::printf("Column of A is not the same size of column of B. Cant compute solution returning A and B unchanged\n");
            return -1;
        }
        const int info = sml_mldivide(h_, A.data(), n > 1 ? n : 1, B.data(), n > 1 ? n : 1, n, nrhs);
        if (info < 0) throw std::runtime_error(std::string("mldivide: ") + sml_last_error(h_));
        if (info != 0) std::printf("something went wrong with dgesv info = %d\nB is not the solution\n", info);
        return info;
    }

    // rolling_average_over_a_period_2d(grid(row0:row0+nrows-1, :), period), src/mod_utilities.f90:1773; grid is
    // (ld, t_len) column-major
    void rolling_average_over_a_period_2d(std::vector<dp> &grid, int ld, int t_len, int period, int row0 = 0, int nrows = -1)
    {
        if (nrows < 0) nrows = ld - row0;
        ck(sml_rolling_average_2d(h_, grid.data() + row0, ld, nrows, t_len, period, 1), "sml_rolling_average_2d");
    }

    void state_get(const reservoir_type &r, std::vector<dp> &x, int kind = ATMO)
    {
        x.resize(r.n);
        ck(sml_state_get(h_, kind, r.assigned_region, x.data()), "sml_state_get");
    }
    void feedback_get(reservoir_type &r, int kind = ATMO)
    {
        r.feedback.resize(r.reservoir_numinputs);
        ck(sml_feedback_get(h_, kind, r.assigned_region, r.feedback.data()), "sml_feedback_get");
    }
    long long kernel_launch_count() const { return sml_kernel_launch_count(h_); }
    sml_engine *handle() { return h_; }

private:
    int ck(int rc, const char *where)
    {
        if (rc < 0) throw std::runtime_error(std::string(where) + ": " + sml_last_error(h_));
        return rc;
    }
    sml_engine *h_ = nullptr;
    int step_predicted_ = -1, ocean_step_predicted_ = -1, current_step_ = 0;
    bool contribs_ = false, fetch_outvecs_ = true;
    int irank_ = 0, numprocs_ = 1;
    std::unordered_map<int, int> local_of_;
    std::vector<dp> outvec_cache_;
};

}  // namespace speedyml
