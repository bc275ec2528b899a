// ncfile.hpp -- minimal reader of NetCDF *classic* files (CDF-1 / CDF-2), host side, no library.
//
// write_trained_res (src/mod_reservoir.f90:1703-1738) stores every trained reservoir in one file created with
// nf90_create(..., NF90_CLOBBER), i.e. the classic format, holding seven fixed-size variables:
//   win(win_y, win_x) float, wout(wout_y, wout_x) float, rows(rows_x) int, cols(cols_x) int, vals(vals_x) float,
//   mean(mean_x) float, std(std_x) float           (src/mod_io.f90:1275-1320: NF90_REAL / NF90_INT, Fortran dimension
// order, so the file's slowest dimension is the Fortran array's last one and the bytes are the column-major array).
// read_trained_res (src/mod_io.f90:2938-2983) reads them back into real(dp): float32 -> FP64 widening happens here.
// Only what that container needs is implemented: non-record variables of type int / float / double.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace sml {

class NcClassicFile {
public:
    struct Var {
        std::vector<int64_t> shape;  // file order (slowest first)
        int type = 0;                // 4 int, 5 float, 6 double
        int64_t begin = 0;
        int64_t count() const
        {
            int64_t c = 1;
            for (int64_t s : shape) c *= s;
            return c;
        }
    };

    explicit NcClassicFile(const std::string &path) : f_(std::fopen(path.c_str(), "rb"))
    {
        if (!f_) throw std::runtime_error("cannot open " + path);
        try {
            if (fseeko(f_, 0, SEEK_END)) throw std::runtime_error("NetCDF seek failed");
            file_size_ = (int64_t)ftello(f_);
            seek(0);
            parse_header();
        } catch (...) {
            std::fclose(f_);
            throw;
        }
    }
    ~NcClassicFile()
    {
        if (f_) std::fclose(f_);
    }
    NcClassicFile(const NcClassicFile &) = delete;
    NcClassicFile &operator=(const NcClassicFile &) = delete;

    bool has(const std::string &name) const { return vars_.count(name) != 0; }
    const Var &var(const std::string &name) const
    {
        auto it = vars_.find(name);
        if (it == vars_.end()) throw std::runtime_error("variable '" + name + "' not in file");
        return it->second;
    }
    // real variable (float or double in the file) widened to FP64, in file byte order of elements
    std::vector<double> read_real(const std::string &name)
    {
        const Var &v = var(name);
        std::vector<double> out((size_t)v.count());
        seek(v.begin);
        if (v.type == 5) {
            std::vector<unsigned char> raw(out.size() * 4);
            rd(raw.data(), raw.size());
            for (size_t i = 0; i < out.size(); ++i) {
                const uint32_t u = (uint32_t)raw[4 * i] << 24 | (uint32_t)raw[4 * i + 1] << 16 | (uint32_t)raw[4 * i + 2] << 8 | raw[4 * i + 3];
                float fl;
                std::memcpy(&fl, &u, 4);
                out[i] = (double)fl;
            }
        } else if (v.type == 6) {
            std::vector<unsigned char> raw(out.size() * 8);
            rd(raw.data(), raw.size());
            for (size_t i = 0; i < out.size(); ++i) {
                uint64_t u = 0;
                for (int b = 0; b < 8; ++b) u = u << 8 | raw[8 * i + b];
                std::memcpy(&out[i], &u, 8);
            }
        } else {
            throw std::runtime_error("variable '" + name + "' is not real");
        }
        return out;
    }
    std::vector<int32_t> read_int(const std::string &name)
    {
        const Var &v = var(name);
        if (v.type != 4) throw std::runtime_error("variable '" + name + "' is not integer");
        std::vector<int32_t> out((size_t)v.count());
        std::vector<unsigned char> raw(out.size() * 4);
        seek(v.begin);
        rd(raw.data(), raw.size());
        for (size_t i = 0; i < out.size(); ++i)
            out[i] = (int32_t)((uint32_t)raw[4 * i] << 24 | (uint32_t)raw[4 * i + 1] << 16 | (uint32_t)raw[4 * i + 2] << 8 | raw[4 * i + 3]);
        return out;
    }

private:
    void rd(void *p, size_t n)
    {
        if (n && std::fread(p, 1, n, f_) != n) throw std::runtime_error("NetCDF file truncated");
    }
    void seek(int64_t off)
    {
        if (fseeko(f_, (off_t)off, SEEK_SET)) throw std::runtime_error("NetCDF seek failed");
    }
    int32_t i32()
    {
        unsigned char b[4];
        rd(b, 4);
        return (int32_t)((uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3]);
    }
    int64_t i64()
    {
        unsigned char b[8];
        rd(b, 8);
        uint64_t u = 0;
        for (int k = 0; k < 8; ++k) u = u << 8 | b[k];
        return (int64_t)u;
    }
    std::string name()
    {
        const int32_t n = i32();
        if (n < 0 || n > 4096) throw std::runtime_error("bad NetCDF name length");
        std::string s((size_t)n, '\0');
        rd(&s[0], (size_t)n);
        skip_pad(n);
        return s;
    }
    void skip_pad(int64_t n)
    {
        const int64_t pad = (4 - n % 4) % 4;
        if (pad && fseeko(f_, (off_t)pad, SEEK_CUR)) throw std::runtime_error("NetCDF seek failed");
    }
    static int type_size(int t)
    {
        switch (t) {
        case 1: case 2: return 1;
        case 3: return 2;
        case 4: case 5: return 4;
        case 6: return 8;
        }
        throw std::runtime_error("unsupported NetCDF type");
    }
    void skip_att_list()
    {
        const int32_t tag = i32(), n = i32();
        if (tag == 0 && n == 0) return;
        if (tag != 0x0C) throw std::runtime_error("bad NetCDF attribute list");
        for (int32_t a = 0; a < n; ++a) {
            name();
            const int32_t t = i32(), ne = i32();
            if (ne < 0) throw std::runtime_error("bad NetCDF attribute length");
            const int64_t bytes = (int64_t)type_size(t) * ne;
            if (bytes > file_size_) throw std::runtime_error("NetCDF attribute runs past the end of the file");
            if (fseeko(f_, (off_t)bytes, SEEK_CUR)) throw std::runtime_error("NetCDF seek failed");
            skip_pad(bytes);
        }
    }
    void parse_header()
    {
        unsigned char magic[4];
        rd(magic, 4);
        if (magic[0] != 'C' || magic[1] != 'D' || magic[2] != 'F' || (magic[3] != 1 && magic[3] != 2))
            throw std::runtime_error("not a NetCDF classic file (HDF5-based NetCDF-4 must be converted: nccopy -k classic)");
        const bool off64 = magic[3] == 2;
        i32();  // numrecs
        std::vector<int64_t> dimlen;
        int32_t tag = i32(), n = i32();
        if (tag == 0x0A) {
            if (n < 0 || n > 1024) throw std::runtime_error("bad NetCDF dimension count");
            for (int32_t d = 0; d < n; ++d) {
                name();
                const int32_t len = i32();
                // length 0 marks the record dimension, which write_trained_res never uses (src/mod_io.f90:1298-1340)
                if (len <= 0) throw std::runtime_error("NetCDF record / non-positive dimensions are not supported");
                dimlen.push_back(len);
            }
        } else if (!(tag == 0 && n == 0)) {
            throw std::runtime_error("bad NetCDF dimension list");
        }
        skip_att_list();
        tag = i32();
        n = i32();
        if (tag == 0 && n == 0) return;
        if (tag != 0x0B) throw std::runtime_error("bad NetCDF variable list");
        if (n < 0 || n > 65536) throw std::runtime_error("bad NetCDF variable count");
        for (int32_t v = 0; v < n; ++v) {
            const std::string nm = name();
            Var var;
            const int32_t nd = i32();
            if (nd < 0 || nd > 1024) throw std::runtime_error("bad NetCDF variable rank");
            for (int32_t d = 0; d < nd; ++d) {
                const int32_t id = i32();
                if (id < 0 || id >= (int32_t)dimlen.size()) throw std::runtime_error("bad NetCDF dimension id");
                var.shape.push_back(dimlen[id]);
            }
            skip_att_list();
            var.type = i32();
            i32();  // vsize
            var.begin = off64 ? i64() : (int64_t)i32();
            // the data block must lie inside the file: count() * element size from begin
            int64_t cnt = 1;
            for (int64_t sdim : var.shape) {   // no overflow: stop as soon as the product leaves the file
                cnt *= sdim;
                if (cnt > file_size_) throw std::runtime_error("NetCDF variable '" + nm + "' is larger than the file");
            }
            if (var.begin < 0 || var.begin > file_size_ ||
                cnt * type_size(var.type) > file_size_ - var.begin)
                throw std::runtime_error("NetCDF variable '" + nm + "' runs past the end of the file");
            vars_[nm] = var;
        }
    }

    std::FILE *f_;
    int64_t file_size_ = 0;
    std::map<std::string, Var> vars_;
};

}  // namespace sml
