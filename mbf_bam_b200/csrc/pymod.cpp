// CPython host extension `mbf_bam_b200.mbf_bam` -- the C++ stand-in for the reference's Rust/PyO3
// cdylib `mbf_bam.mbf_bam` (reference src/lib.rs).  It only converts Python objects to the flat
// arrays of include/mbfbam.h, calls the C ABI of libmbfbam_cuda.so, and assembles the result
// dicts.  No alignment data is touched here and there is no CPU counting path: without the CUDA
// library / a GPU every call raises.
//
//   count_reads_unstranded   src/lib.rs:138-157 + map assembly of counters.rs:236-254
//   count_reads_stranded     src/lib.rs:161-181 + counters.rs:262-273,309-314
//   count_introns            src/lib.rs:216-225
//   count_reads_weighted     not in the reference
//   quantify_gene_reads      src/lib.rs:247-264 + quantify.rs:20-75
//   calculate_duplicate_distribution   src/lib.rs:108-116
//   count_reads_primary_only_right_strand_only_by_barcode   src/lib.rs:184-211 + by_barcode.rs:18-92
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/mbfbam.h"

namespace py = pybind11;

namespace {

struct Flat {
    std::vector<std::string> names;
    std::vector<const char*> name_ptrs;
    std::vector<uint32_t> n_genes;
    std::vector<uint64_t> offs{0};
    std::vector<uint32_t> st, en, ge;
    std::vector<int8_t> sd;
    // the caller's gene id objects (strong references: the GIL is released while the GPU works); they
    // become the keys of the result dicts, so no string is copied or hashed again
    std::vector<std::vector<PyObject*>> gene_ids;
    mbfbam_intervals c{};
    Flat() = default;
    Flat(const Flat&) = delete;
    Flat& operator=(const Flat&) = delete;
    ~Flat() {
        for (auto& v : gene_ids)
            for (PyObject* o : v) Py_DECREF(o);
    }
};

[[noreturn]] void value_error(const std::string& m) {
    PyErr_SetString(PyExc_ValueError, m.c_str());
    throw py::error_already_set();
}

// src/lib.rs:118-134 py_intervals_to_trees + src/count_reads/mod.rs:27-45 build_tree
void flatten(const py::dict& d, Flat& f) {
    // plain CPython calls: this runs once per count call over ~10^5 small objects, and the generic
    // pybind11 casts cost about a microsecond each
    auto as_ll = [](PyObject* o, long long& out) {
        out = PyLong_AsLongLong(o);
        if (out == -1 && PyErr_Occurred()) {
            PyErr_Clear();
            value_error("Python error: interval coordinate/strand is not an integer");
        }
    };
    PyObject *key, *value;
    Py_ssize_t pos = 0;
    while (PyDict_Next(d.ptr(), &pos, &key, &value)) {
        Py_ssize_t klen = 0;
        const char* kp = PyUnicode_Check(key) ? PyUnicode_AsUTF8AndSize(key, &klen) : nullptr;
        if (!kp) {
            PyErr_Clear();
            value_error("Python error: interval key is not a str");
        }
        f.names.emplace_back(kp, (size_t)klen);
        if (!PyList_Check(value)) value_error("Python error: interval value is not a list");
        const Py_ssize_t ng = PyList_GET_SIZE(value);
        f.n_genes.push_back((uint32_t)ng);
        f.gene_ids.emplace_back();
        auto& ids = f.gene_ids.back();
        ids.reserve((size_t)ng);
        for (Py_ssize_t g = 0; g < ng; g++) {
            PyObject* t = PyList_GET_ITEM(value, g);
            if (!PyTuple_Check(t)) value_error("Python error: interval entry is not a tuple");
            if (PyTuple_GET_SIZE(t) < 4) value_error("Python error: interval entry needs (gene_id, strand, starts, stops)");
            PyObject* id = PyTuple_GET_ITEM(t, 0);
            if (!PyUnicode_Check(id)) value_error("Python error: gene id is not a str");
            Py_INCREF(id);
            ids.push_back(id);
            long long strand;
            as_ll(PyTuple_GET_ITEM(t, 1), strand);
            if (strand < -128 || strand > 127) value_error("Python error: strand does not fit i8");
            PyObject* s = PyTuple_GET_ITEM(t, 2);
            PyObject* e = PyTuple_GET_ITEM(t, 3);
            if (!PyList_Check(s) || !PyList_Check(e)) value_error("Python error: starts/stops must be lists");
            const Py_ssize_t n = std::min(PyList_GET_SIZE(s), PyList_GET_SIZE(e));  // zip() truncates (mod.rs:38)
            for (Py_ssize_t i = 0; i < n; i++) {
                long long a, b;
                as_ll(PyList_GET_ITEM(s, i), a);
                as_ll(PyList_GET_ITEM(e, i), b);
                if (a < 0 || b < 0 || a > 0xffffffffLL || b > 0xffffffffLL) value_error("Python error: coordinate does not fit u32");
                f.st.push_back((uint32_t)a);
                f.en.push_back((uint32_t)b);
                f.ge.push_back((uint32_t)g);
                f.sd.push_back((int8_t)strand);
            }
        }
        f.offs.push_back(f.st.size());
    }
    for (auto& n : f.names) f.name_ptrs.push_back(n.c_str());
    f.c.n_contigs = (int32_t)f.names.size();
    f.c.contig_names = f.name_ptrs.data();
    f.c.n_genes = f.n_genes.data();
    f.c.iv_offsets = f.offs.data();
    f.c.iv_start = f.st.data();
    f.c.iv_end = f.en.data();
    f.c.iv_gene = f.ge.data();
    f.c.iv_strand = f.sd.data();
}

// result dict helpers (keys are the caller's own str objects or small "_name" strings)
inline PyObject* to_py(uint64_t v) { return PyLong_FromUnsignedLongLong(v); }
inline PyObject* to_py(double v) { return PyFloat_FromDouble(v); }
inline bool from_py(PyObject* o, uint64_t& v) { v = PyLong_AsUnsignedLongLong(o); return !(v == (uint64_t)-1 && PyErr_Occurred()); }
inline bool from_py(PyObject* o, double& v) { v = PyFloat_AsDouble(o); return !(v == -1.0 && PyErr_Occurred()); }

template <class T>
void dict_add(PyObject* d, PyObject* key, T v) {  // d[key] = d.get(key, 0) + v
    PyObject* nv = to_py(v);
    if (!nv) throw py::error_already_set();
    // one lookup inserts the value when the key is new (the usual case: gene ids are unique)
    PyObject* cur = PyDict_SetDefault(d, key, nv);  // borrowed
    if (!cur) {
        Py_DECREF(nv);
        throw py::error_already_set();
    }
    if (cur == nv) {
        Py_DECREF(nv);
        return;
    }
    Py_DECREF(nv);
    T o;
    if (!from_py(cur, o)) throw py::error_already_set();
    nv = to_py((T)(v + o));
    if (!nv || PyDict_SetItem(d, key, nv) != 0) {
        Py_XDECREF(nv);
        throw py::error_already_set();
    }
    Py_DECREF(nv);
}
template <class T>
void dict_add(PyObject* d, const std::string& key, T v) {
    py::str k(key);
    dict_add<T>(d, k.ptr(), v);
}


py::dict stats_dict(const mbfbam_stats& s) {
    py::dict d;
    d["n_records"] = s.n_records;
    d["n_records_owned"] = s.n_records_owned;
    d["n_bgzf_blocks"] = s.n_bgzf_blocks;
    d["compressed_bytes"] = s.compressed_bytes;
    d["uncompressed_bytes"] = s.uncompressed_bytes;
    d["h2d_bytes"] = s.h2d_bytes;
    d["d2h_bytes"] = s.d2h_bytes;
    d["n_chunks"] = s.n_chunks;
    d["n_kernel_launches"] = s.n_kernel_launches;
    d["n_windows"] = s.n_windows;
    d["n_mm_keys"] = s.n_mm_keys;
    d["n_fixups"] = s.n_fixups;
    d["ms_total"] = s.ms_total;
    d["ms_gpu_scan"] = s.ms_gpu_scan;
    d["ms_gpu_inflate"] = s.ms_gpu_inflate;
    d["ms_gpu_frame"] = s.ms_gpu_frame;
    d["ms_gpu_count"] = s.ms_gpu_count;
    d["ms_gpu_other"] = s.ms_gpu_other;
    d["ms_gpu_tokens"] = s.ms_gpu_tokens;
    d["n_tokens"] = s.n_tokens;
    d["ms_gpu_all"] = s.ms_gpu_all;
    d["h2d_pinned_bytes"] = s.h2d_pinned_bytes;
    d["n_devices"] = s.n_devices;
    d["n_retries"] = s.n_retries;
    d["ms_pin"] = s.ms_pin;
    return d;
}

// Devices a call fans out over when the caller did not say (Reader.set_devices): MBF_BAM_DEVICES ("all" or a
// comma list); else, inside a torchrun/torch.distributed job (LOCAL_RANK set), this rank's GPU only; else every
// visible GPU, but no more than one per 256 MiB of file (a small file is not worth eight pipelines).
std::vector<int32_t> default_devices(uint64_t file_bytes) {
    std::vector<int32_t> d;
    const int n = mbfbam_device_count();
    const char* v = getenv("MBF_BAM_DEVICES");
    if (v && *v && strcmp(v, "all") != 0) {
        for (const char* p = v; *p;) {
            char* e = nullptr;
            long x = strtol(p, &e, 10);
            if (e == p) break;
            d.push_back((int32_t)x);
            p = *e == ',' ? e + 1 : e;
        }
        if (!d.empty()) return d;
    }
    const char* lr = getenv("LOCAL_RANK");
    if (!(v && *v) && lr && *lr) {
        d.push_back((int32_t)atoi(lr));
        return d;
    }
    const uint64_t per = 256ull << 20;
    int want = (int)std::min<uint64_t>((uint64_t)std::max(1, n), std::max<uint64_t>(1, (file_bytes + per - 1) / per));
    if (v && !strcmp(v, "all")) want = std::max(1, n);
    for (int i = 0; i < want; i++) d.push_back(i);
    return d;
}

struct PyReader {
    mbfbam_reader* r = nullptr;
    mbfbam_stats last{};
    std::vector<int32_t> devices;
    int shard_index = 0, shard_count = 1;
    uint64_t file_bytes = 0;

    PyReader(const std::string& filename, py::object index_filename) {
        char err[512] = {0};
        std::string idx;
        const char* ip = nullptr;
        if (!index_filename.is_none()) {
            idx = py::cast<std::string>(index_filename);
            ip = idx.c_str();
        }
        if (mbfbam_open(filename.c_str(), ip, &r, err, sizeof err)) value_error(err);
        file_bytes = mbfbam_file_size(r);
        devices = default_devices(file_bytes);
    }
    ~PyReader() { mbfbam_close(r); }

    void set_device(int d) { devices.assign(1, d); }
    void set_devices(std::vector<int32_t> d) {
        if (d.empty()) value_error("set_devices: empty device list");
        devices = std::move(d);
    }
    std::vector<int32_t> get_devices() const { return devices; }
    // explicit sharding for one-process-per-GPU jobs (mbf_bam_b200.dist); never read from the environment
    void set_shard(int i, int n) {
        if (n < 1 || i < 0 || i >= n) value_error("set_shard: need 0 <= index < count");
        shard_index = i;
        shard_count = n;
    }

    void preload() {
        char err[512] = {0};
        int rc;
        {
            py::gil_scoped_release rel;
            rc = mbfbam_preload(r, devices[0], err, sizeof err);
        }
        if (rc) value_error(err);
    }
    void drop_preload() { mbfbam_drop_preload(r); }
    bool pin_file() {
        char err[512] = {0};
        int rc;
        {
            py::gil_scoped_release rel;
            rc = mbfbam_pin_file(r, err, sizeof err);
        }
        if (rc == MBFBAM_ERR_UNSUPPORTED) {
            if (getenv("MBF_BAM_TRACE")) fprintf(stderr, "[mbfbam] pin_file: %s\n", err);
            return false;
        }
        if (rc) value_error(err);
        return true;
    }
    void unpin_file() { mbfbam_unpin_file(r); }

    std::vector<std::pair<std::string, uint32_t>> targets() const {
        std::vector<std::pair<std::string, uint32_t>> v;
        for (int32_t i = 0; i < mbfbam_n_targets(r); i++) v.push_back({mbfbam_target_name(r, i), mbfbam_target_len(r, i)});
        return v;
    }

    py::list chunks(py::object gene_intervals) {
        Flat g;
        const mbfbam_intervals* gp = nullptr;
        if (!gene_intervals.is_none()) {
            flatten(py::cast<py::dict>(gene_intervals), g);
            gp = &g.c;
        }
        int64_t n = mbfbam_chunks(r, gp, nullptr, nullptr, nullptr, 0);
        if (n < 0) value_error("could not build the chunk table");
        std::vector<int32_t> tid((size_t)n + 1);
        std::vector<uint32_t> st((size_t)n + 1), en((size_t)n + 1);
        mbfbam_chunks(r, gp, tid.data(), st.data(), en.data(), n);
        py::list out;
        for (int64_t i = 0; i < n; i++) out.append(py::make_tuple(tid[(size_t)i], st[(size_t)i], en[(size_t)i]));
        return out;
    }

    py::tuple plan(py::object gene_intervals, int shard, int nshards) {
        Flat g;
        const mbfbam_intervals* gp = nullptr;
        if (!gene_intervals.is_none()) {
            flatten(py::cast<py::dict>(gene_intervals), g);
            gp = &g.c;
        }
        int64_t lo = 0, hi = 0;
        std::vector<uint64_t> b(4096), e(4096);
        int64_t n = mbfbam_plan_shard(r, gp, shard, nshards, &lo, &hi, b.data(), e.data(), 4096);
        if (n < 0) value_error("could not plan the shard");
        py::list ranges;
        for (int64_t i = 0; i < n && i < 4096; i++) ranges.append(py::make_tuple(b[(size_t)i], e[(size_t)i]));
        return py::make_tuple(lo, hi, ranges);
    }

    // dense results: used by the dict assembly below and by the multi-process (one rank per GPU) merge
    struct Raw {
        Flat iv;
        std::vector<uint64_t> fwd, rev;
        std::vector<double> wfwd, wrev;
        std::vector<uint8_t> scanned;
        uint64_t outside = 0;
    };
    std::unique_ptr<Raw> kept;       // count_reads_raw -> assemble hand-over
    py::object kept_for = py::none();

    void run(int mode, const py::dict& intervals, const py::dict& gene_intervals, bool once, Raw& raw) {
        Flat gv;
        flatten(intervals, raw.iv);
        flatten(gene_intervals, gv);
        size_t G = 0;
        for (uint32_t n : raw.iv.n_genes) G += n;
        raw.fwd.assign(G + 1, 0);
        raw.rev.assign(G + 1, 0);
        raw.wfwd.assign(G + 1, 0.0);
        raw.wrev.assign(G + 1, 0.0);
        raw.scanned.assign(raw.iv.names.size() + 1, 0);
        mbfbam_count_job job{};
        job.intervals = &raw.iv.c;
        job.gene_intervals = &gv.c;
        job.mode = mode;
        job.each_read_counts_once = once ? 1 : 0;
        job.device = devices[0];
        job.n_devices = (int32_t)devices.size();
        job.devices = devices.data();
        job.shard_index = shard_index;
        job.shard_count = shard_count;
        mbfbam_counts out{};
        out.fwd = raw.fwd.data();
        out.rev = raw.rev.data();
        out.wfwd = raw.wfwd.data();
        out.wrev = raw.wrev.data();
        out.scanned = raw.scanned.data();
        char err[512] = {0};
        int rc;
        {
            py::gil_scoped_release rel;  // superset of the reference, which keeps the GIL
            rc = mbfbam_count_reads(r, &job, &out, &last, err, sizeof err);
        }
        if (rc) value_error(err);
        raw.outside = out.outside;
    }

    // counters.rs:236-254 summed over chunks: per contig, `insert` keeps the LAST gene_no of a duplicated id
    static py::dict assemble_unstranded(const Raw& raw) {
        py::dict out;
        size_t base = 0;
        uint64_t total_all = 0;
        bool any = false;
        for (size_t c = 0; c < raw.iv.names.size(); c++) {
            const auto& ids = raw.iv.gene_ids[c];
            if (raw.scanned[c]) {
                any = true;
                py::dict last;  // insert semantics within one contig
                uint64_t tot = 0;
                for (size_t g = 0; g < ids.size(); g++) {
                    PyObject* v = to_py(raw.fwd[base + g]);
                    if (!v || PyDict_SetItem(last.ptr(), ids[g], v) != 0) {
                        Py_XDECREF(v);
                        throw py::error_already_set();
                    }
                    Py_DECREF(v);
                    tot += raw.fwd[base + g];
                }
                PyObject *k, *v;
                Py_ssize_t pos = 0;
                while (PyDict_Next(last.ptr(), &pos, &k, &v)) {
                    uint64_t x;
                    if (!from_py(v, x)) throw py::error_already_set();
                    dict_add<uint64_t>(out.ptr(), k, x);
                }
                dict_add<uint64_t>(out.ptr(), "_" + raw.iv.names[c], tot);
                total_all += tot;
            }
            base += ids.size();
        }
        if (any) {
            dict_add<uint64_t>(out.ptr(), std::string("_total"), total_all);
            dict_add<uint64_t>(out.ptr(), std::string("_outside"), raw.outside);
        }
        return out;
    }

    // counters.rs:262-273 (+= per gene id), :309-314 (_outside only in the forward dict)
    template <class T>
    static py::tuple assemble_dual(const Raw& raw, const std::vector<T>& f, const std::vector<T>& v) {
        py::dict da, db;
        size_t base = 0;
        bool any = false;
        T all_a = 0, all_b = 0;
        for (size_t c = 0; c < raw.iv.names.size(); c++) {
            const auto& ids = raw.iv.gene_ids[c];
            if (raw.scanned[c]) {
                any = true;
                T ta = 0, tb = 0;
                for (size_t g = 0; g < ids.size(); g++) {
                    dict_add<T>(da.ptr(), ids[g], f[base + g]);
                    dict_add<T>(db.ptr(), ids[g], v[base + g]);
                    ta += f[base + g];
                    tb += v[base + g];
                }
                dict_add<T>(da.ptr(), "_" + raw.iv.names[c], ta);
                dict_add<T>(db.ptr(), "_" + raw.iv.names[c], tb);
                all_a += ta;
                all_b += tb;
            }
            base += ids.size();
        }
        if (any) {
            dict_add<T>(da.ptr(), std::string("_total"), all_a);
            dict_add<T>(db.ptr(), std::string("_total"), all_b);
            da["_outside"] = py::int_(raw.outside);
        }
        return py::make_tuple(da, db);
    }

    py::object count_reads(int mode, const py::dict& intervals, const py::dict& gene_intervals, py::object once) {
        Raw raw;
        run(mode, intervals, gene_intervals, !once.is_none() && py::cast<bool>(once), raw);
        if (mode == MBFBAM_MODE_UNSTRANDED) return assemble_unstranded(raw);
        if (mode == MBFBAM_MODE_STRANDED) return assemble_dual<uint64_t>(raw, raw.fwd, raw.rev);
        return assemble_dual<double>(raw, raw.wfwd, raw.wrev);
    }

    // dense vectors (bytes of u64 / f64) for the one-process-per-GPU merge: the caller all-reduces them
    // and calls assemble()
    py::tuple count_reads_raw(int mode, const py::dict& intervals, const py::dict& gene_intervals, py::object once) {
        std::unique_ptr<Raw> rawp(new Raw());
        Raw& raw = *rawp;
        run(mode, intervals, gene_intervals, !once.is_none() && py::cast<bool>(once), raw);
        size_t G = raw.fwd.size() - 1;
        auto b = [](const void* p, size_t n) { return py::bytes((const char*)p, n); };
        py::tuple out = py::make_tuple(b(raw.fwd.data(), G * 8), b(raw.rev.data(), G * 8), b(raw.wfwd.data(), G * 8),
                                       b(raw.wrev.data(), G * 8), b(raw.scanned.data(), raw.iv.names.size()), raw.outside);
        // assemble() for the same `intervals` object follows at once (mbf_bam_b200.dist): keep the flattened form
        kept = std::move(rawp);
        kept_for = py::reinterpret_borrow<py::object>(intervals);
        return out;
    }

    py::object assemble(int mode, const py::dict& intervals, py::bytes fwd, py::bytes rev, py::bytes wfwd, py::bytes wrev,
                        py::bytes scanned, uint64_t outside) {
        std::unique_ptr<Raw> rawp;
        if (kept && kept_for.ptr() == intervals.ptr()) {
            rawp = std::move(kept);
        } else {
            rawp.reset(new Raw());
            flatten(intervals, rawp->iv);
        }
        kept.reset();
        kept_for = py::none();
        Raw& raw = *rawp;
        size_t G = 0;
        for (uint32_t n : raw.iv.n_genes) G += n;
        auto load = [&](py::bytes& src, void* dst, size_t n) {
            std::string s = src;
            if (s.size() != n) value_error("assemble: wrong buffer size");
            memcpy(dst, s.data(), n);
        };
        raw.fwd.assign(G + 1, 0); raw.rev.assign(G + 1, 0); raw.wfwd.assign(G + 1, 0); raw.wrev.assign(G + 1, 0);
        raw.scanned.assign(raw.iv.names.size() + 1, 0);
        load(fwd, raw.fwd.data(), G * 8); load(rev, raw.rev.data(), G * 8);
        load(wfwd, raw.wfwd.data(), G * 8); load(wrev, raw.wrev.data(), G * 8);
        load(scanned, raw.scanned.data(), raw.iv.names.size());
        raw.outside = outside;
        if (mode == MBFBAM_MODE_UNSTRANDED) return assemble_unstranded(raw);
        if (mode == MBFBAM_MODE_STRANDED) return assemble_dual<uint64_t>(raw, raw.fwd, raw.rev);
        return assemble_dual<double>(raw, raw.wfwd, raw.wrev);
    }

    py::dict count_introns() {
        mbfbam_introns res{};
        char err[512] = {0};
        int rc;
        {
            py::gil_scoped_release rel;
            rc = mbfbam_count_introns_multi(r, devices.data(), (int32_t)devices.size(), shard_index, shard_count, &res, &last, err, sizeof err);
        }
        if (rc) {
            mbfbam_free_introns(&res);
            value_error(err);
        }
        py::dict out;
        std::vector<py::dict> per(res.n_targets > 0 ? (size_t)res.n_targets : 0);
        for (int32_t t = 0; t < res.n_targets; t++)
            if (res.contig_seen[t]) out[py::str(mbfbam_target_name(r, t))] = per[(size_t)t];
        // every (tid, start, stop) appears once in the result (include/mbfbam.h): plain inserts, plain CPython calls --
        // a whole-transcriptome file yields 10^5..10^7 keys and this loop is then most of the call
        bool failed = false;
        for (uint64_t i = 0; i < res.n && !failed; i++) {
            py::dict& d = per[(size_t)res.tid[i]];
            PyObject* a = PyLong_FromUnsignedLong(res.start[i]);
            PyObject* b = PyLong_FromUnsignedLong(res.stop[i]);
            PyObject* v = PyLong_FromUnsignedLongLong(res.count[i]);
            PyObject* key = (a && b) ? PyTuple_Pack(2, a, b) : nullptr;
            if (!key || !v || PyDict_SetItem(d.ptr(), key, v) != 0) failed = true;
            Py_XDECREF(a); Py_XDECREF(b); Py_XDECREF(v); Py_XDECREF(key);
            if (!failed && !res.contig_seen[res.tid[i]]) out[py::str(mbfbam_target_name(r, res.tid[i]))] = d;
        }
        if (failed) {
            mbfbam_free_introns(&res);
            throw py::error_already_set();
        }
        mbfbam_free_introns(&res);
        return out;
    }

    // quantify.rs:20-75: (congruent, incongruent) {gene_id: [(pos, count)]}.  Every gene of a scanned contig gets a
    // list (to_hashmap, quantify.rs:148-157: within one contig the LAST gene_no of a duplicated id wins); equal ids
    // on different contigs are merged and a (gene_id, pos) seen twice is the reference's "still not disjoint" panic
    // (quantify.rs:189-203) -> ValueError.  Lists are sorted by position (HashMap order in the reference).
    py::tuple quantify_gene_reads(const py::dict& intervals, const py::dict& gene_intervals) {
        Flat iv, gv;
        flatten(intervals, iv);
        flatten(gene_intervals, gv);
        mbfbam_count_job job{};
        job.intervals = &iv.c;
        job.gene_intervals = &gv.c;
        job.mode = MBFBAM_MODE_STRANDED;
        job.device = devices[0];
        job.n_devices = (int32_t)devices.size();
        job.devices = devices.data();
        job.shard_index = shard_index;
        job.shard_count = shard_count;
        mbfbam_gene_pos_counts res{};
        char err[512] = {0};
        int rc;
        {
            py::gil_scoped_release rel;
            rc = mbfbam_quantify_gene_reads(r, &job, &res, &last, err, sizeof err);
        }
        if (rc) {
            mbfbam_free_gene_pos_counts(&res);
            value_error(err);
        }
        size_t G = 0;
        for (uint32_t n : iv.n_genes) G += n;
        // entries grouped by (bucket, gene), sorted by position
        std::vector<uint64_t> order(res.n);
        for (uint64_t i = 0; i < res.n; i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) {
            if (res.bucket[a] != res.bucket[b]) return res.bucket[a] < res.bucket[b];
            if (res.gene[a] != res.gene[b]) return res.gene[a] < res.gene[b];
            return res.pos[a] < res.pos[b];
        });
        std::vector<uint64_t> first[2];
        first[0].assign(G + 1, 0);
        first[1].assign(G + 1, 0);
        for (uint64_t i = 0; i < res.n; i++)
            if (res.gene[i] < G) first[res.bucket[i] & 1][res.gene[i] + 1]++;
        uint64_t n0 = 0;
        for (size_t g = 0; g < G; g++) n0 += first[0][g + 1];
        for (int b = 0; b < 2; b++) {
            first[b][0] = b ? n0 : 0;
            for (size_t g = 0; g < G; g++) first[b][g + 1] += first[b][g];
        }
        py::dict out[2];
        bool fail = false;
        for (int b = 0; b < 2 && !fail; b++) {
            size_t base = 0;
            for (size_t c = 0; c < iv.names.size() && !fail; c++) {
                const auto& ids = iv.gene_ids[c];
                if (res.scanned[c]) {
                    py::dict last_of;  // id -> last gene_no of this contig
                    for (size_t g = 0; g < ids.size(); g++) {
                        PyObject* v = PyLong_FromSize_t(g);
                        if (!v || PyDict_SetItem(last_of.ptr(), ids[g], v) != 0) {
                            Py_XDECREF(v);
                            mbfbam_free_gene_pos_counts(&res);
                            throw py::error_already_set();
                        }
                        Py_DECREF(v);
                    }
                    PyObject *k, *v;
                    Py_ssize_t pos = 0;
                    while (PyDict_Next(last_of.ptr(), &pos, &k, &v)) {
                        const size_t g = base + PyLong_AsSize_t(v);
                        py::list lst;
                        for (uint64_t i = first[b][g]; i < first[b][g + 1]; i++)
                            lst.append(py::make_tuple(res.pos[order[i]], res.count[order[i]]));
                        PyObject* old = PyDict_GetItemWithError(out[b].ptr(), k);
                        if (old) {  // the same id on an earlier contig
                            py::set seen;
                            py::list ol = py::reinterpret_borrow<py::list>(old);
                            for (auto t : ol) seen.add(t.cast<py::tuple>()[0]);
                            for (auto t : lst) {
                                if (seen.contains(t.cast<py::tuple>()[0])) fail = true;
                                ol.append(t);
                            }
                            ol.attr("sort")();
                        } else if (PyDict_SetItem(out[b].ptr(), k, lst.ptr()) != 0) {
                            mbfbam_free_gene_pos_counts(&res);
                            throw py::error_already_set();
                        }
                    }
                }
                base += ids.size();
            }
        }
        mbfbam_free_gene_pos_counts(&res);
        if (fail) value_error("still not disjoint");
        return py::make_tuple(out[0], out[1]);
    }

    // by_barcode.rs:18-92: (genes, barcodes, [(gene_idx, barcode_idx, count)]), one entry per (chunk, gene_no, barcode).
    // Indices are handed out in order of first appearance; the reference walks HashMaps, so its order is arbitrary --
    // here: chunks in order, then gene slot, then barcode bytes.
    py::tuple count_by_barcode(const py::dict& intervals, const py::dict& gene_intervals, const std::string& umi_strategy) {
        if (umi_strategy != "straight") value_error("invalid umi_strategy");  // src/lib.rs:198-201
        Flat iv, gv;
        flatten(intervals, iv);
        flatten(gene_intervals, gv);
        mbfbam_count_job job{};
        job.intervals = &iv.c;
        job.gene_intervals = &gv.c;
        job.mode = MBFBAM_MODE_STRANDED;
        job.device = devices[0];
        job.n_devices = (int32_t)devices.size();
        job.devices = devices.data();
        job.shard_index = shard_index;
        job.shard_count = shard_count;
        mbfbam_barcode_counts res{};
        char err[512] = {0};
        int rc;
        {
            py::gil_scoped_release rel;
            rc = mbfbam_count_by_barcode(r, &job, &res, &last, err, sizeof err);
        }
        if (rc) {
            mbfbam_free_barcode_counts(&res);
            value_error(err);
        }
        std::vector<PyObject*> slot_id;  // global gene slot -> the caller's gene id object (borrowed)
        for (const auto& ids : iv.gene_ids)
            for (PyObject* o : ids) slot_id.push_back(o);
        py::dict gene_idx, bc_idx;
        py::list genes, barcodes, counts;
        try {
            for (uint64_t i = 0; i < res.n; i++) {
                if (res.gene[i] >= slot_id.size()) value_error("gene slot out of range");
                py::object g = py::reinterpret_borrow<py::object>(slot_id[res.gene[i]]);
                py::object b = py::reinterpret_steal<py::object>(
                    PyUnicode_DecodeUTF8(res.bc_bytes + res.bc_off[i], (Py_ssize_t)(res.bc_off[i + 1] - res.bc_off[i]), nullptr));
                if (!b) {
                    PyErr_Clear();
                    value_error("XC barcode is not valid UTF-8");  // the reference unwraps from_utf8 (by_barcode.rs:57)
                }
                if (!gene_idx.contains(g)) { gene_idx[g] = py::int_(genes.size()); genes.append(g); }
                if (!bc_idx.contains(b)) { bc_idx[b] = py::int_(barcodes.size()); barcodes.append(b); }
                counts.append(py::make_tuple(gene_idx[g], bc_idx[b], res.count[i]));
            }
        } catch (...) {
            mbfbam_free_barcode_counts(&res);
            throw;
        }
        mbfbam_free_barcode_counts(&res);
        return py::make_tuple(genes, barcodes, counts);
    }

    // duplicate_distribution.rs:44-86 -> {k: occurrences}
    py::dict duplicate_distribution() {
        mbfbam_dup_hist res{};
        char err[512] = {0};
        int rc;
        {
            py::gil_scoped_release rel;
            rc = mbfbam_duplicate_distribution(r, devices.data(), (int32_t)devices.size(), shard_index, shard_count, &res, &last, err,
                                               sizeof err);
        }
        if (rc) {
            mbfbam_free_dup_hist(&res);
            value_error(err);
        }
        py::dict out;
        for (uint64_t i = 0; i < res.n; i++) out[py::int_(res.k[i])] = py::int_(res.count[i]);
        mbfbam_free_dup_hist(&res);
        return out;
    }

    py::dict stats() const { return stats_dict(last); }
};

mbfbam_stats g_last_stats{};  // POD: a global py::object would be destroyed after the interpreter

py::object module_count(int mode, const std::string& filename, py::object index_filename, const py::dict& intervals,
                        const py::dict& gene_intervals, py::object once) {
    PyReader rd(filename, index_filename);
    py::object res = rd.count_reads(mode, intervals, gene_intervals, once);
    g_last_stats = rd.last;
    return res;
}

}  // namespace

PYBIND11_MODULE(mbf_bam, m) {
    m.doc() = "B200-native drop-in for the read counters of mbf_bam.mbf_bam (reference src/lib.rs:286-299)";
    m.attr("__version__") = mbfbam_version();

    m.def("count_reads_unstranded",
          [](const std::string& filename, py::object index_filename, const py::dict& intervals, const py::dict& gene_intervals,
             py::object each_read_counts_once) {
              return module_count(MBFBAM_MODE_UNSTRANDED, filename, index_filename, intervals, gene_intervals, each_read_counts_once);
          },
          py::arg("filename"), py::arg("index_filename"), py::arg("intervals"), py::arg("gene_intervals"),
          py::arg("each_read_counts_once") = py::none(), "python wrapper for py_count_reads_unstranded (reference src/lib.rs:138)");
    m.def("count_reads_stranded",
          [](const std::string& filename, py::object index_filename, const py::dict& intervals, const py::dict& gene_intervals,
             py::object each_read_counts_once) {
              return module_count(MBFBAM_MODE_STRANDED, filename, index_filename, intervals, gene_intervals, each_read_counts_once);
          },
          py::arg("filename"), py::arg("index_filename"), py::arg("intervals"), py::arg("gene_intervals"),
          py::arg("each_read_counts_once") = py::none(), "python wrapper for py_count_reads_stranded (reference src/lib.rs:161)");
    m.def("count_reads_weighted",
          [](const std::string& filename, py::object index_filename, const py::dict& intervals, const py::dict& gene_intervals,
             py::object each_read_counts_once) {
              return module_count(MBFBAM_MODE_WEIGHTED, filename, index_filename, intervals, gene_intervals, each_read_counts_once);
          },
          py::arg("filename"), py::arg("index_filename"), py::arg("intervals"), py::arg("gene_intervals"),
          py::arg("each_read_counts_once") = py::none(), "1/NH weighted stranded counts (not in the reference; DESIGN.md)");
    m.def("count_introns",
          [](const std::string& filename, py::object index_filename) {
              PyReader rd(filename, index_filename);
              py::dict res = rd.count_introns();
              g_last_stats = rd.last;
              return res;
          },
          py::arg("filename"), py::arg("index_filename") = py::none(), "python wrapper for py_count_introns (reference src/lib.rs:216)");
    m.def("quantify_gene_reads",
          [](const std::string& filename, py::object index_filename, const py::dict& intervals, const py::dict& gene_intervals) {
              PyReader rd(filename, index_filename);
              py::tuple res = rd.quantify_gene_reads(intervals, gene_intervals);
              g_last_stats = rd.last;
              return res;
          },
          py::arg("filename"), py::arg("index_filename"), py::arg("intervals"), py::arg("gene_intervals"),
          "python wrapper for py_quantify_gene_reads (reference src/lib.rs:247)");
    m.def("calculate_duplicate_distribution",
          [](const std::string& filename, py::object index_filename) {
              PyReader rd(filename, index_filename);
              py::dict res = rd.duplicate_distribution();
              g_last_stats = rd.last;
              return res;
          },
          py::arg("filename"), py::arg("index_filename") = py::none(),
          "python wrapper for py_calculate_duplicate_distribution (reference src/lib.rs:108)");
    m.def("count_reads_primary_only_right_strand_only_by_barcode",
          [](const std::string& filename, py::object index_filename, const py::dict& intervals, const py::dict& gene_intervals,
             const std::string& umi_strategy) {
              PyReader rd(filename, index_filename);
              py::tuple res = rd.count_by_barcode(intervals, gene_intervals, umi_strategy);
              g_last_stats = rd.last;
              return res;
          },
          py::arg("filename"), py::arg("index_filename"), py::arg("intervals"), py::arg("gene_intervals"), py::arg("umi_strategy"),
          "python wrapper for py_count_reads_primary_only_right_strand_only_by_barcode (reference src/lib.rs:184)");
    m.def("last_stats", []() { return stats_dict(g_last_stats); }, "measurements of the most recent module-level call");
    m.def("flatten_stats", [](const py::dict& intervals) {
        Flat f;
        flatten(intervals, f);
        size_t g = 0;
        for (uint32_t n : f.n_genes) g += n;
        return py::make_tuple(f.names.size(), g, f.st.size());
    }, py::arg("intervals"), "(contigs, genes, intervals) of an interval dict after the conversion every count call does");
    m.def("device_count", []() { return mbfbam_device_count(); });
    m.def("release_device", [](int device) { mbfbam_release_device(device); }, py::arg("device") = 0,
          "free the cached pipeline (device + pinned buffers) of a device");
    m.def("host_cache_clear", []() { mbfbam_host_cache_clear(); }, "drop the process-wide file mappings and their page locks");

    m.def("inflate_buffer", [](py::bytes data, int device) {
        std::string s = data;
        uint64_t need = 0;
        // sum of ISIZE by walking the BSIZE chain (host, headers only)
        for (size_t off = 0; off + 28 <= s.size();) {
            uint32_t bs = ((uint8_t)s[off + 16] | ((uint8_t)s[off + 17] << 8)) + 1u;
            if (off + bs > s.size()) break;
            uint32_t isz;
            memcpy(&isz, s.data() + off + bs - 4, 4);
            need += isz;
            off += bs;
        }
        std::string out(need, '\0');
        uint64_t got = 0;
        char err[512] = {0};
        mbfbam_stats st{};
        int rc;
        {
            py::gil_scoped_release rel;
            rc = mbfbam_inflate_buffer((const uint8_t*)s.data(), s.size(), (uint8_t*)&out[0], need, &got, device, &st, err, sizeof err);
        }
        if (rc) value_error(err);
        out.resize(got);
        g_last_stats = st;
        return py::bytes(out);
    }, py::arg("data"), py::arg("device") = 0, "inflate every BGZF member of `data` on the GPU (stage-level parity entry)");

    m.def("cigar_expand", [](std::vector<uint32_t> cigar, int pos, bool introns, int device) {
        std::vector<uint32_t> out(2 * (cigar.size() + 1));
        int32_t n = 0;
        char err[512] = {0};
        if (mbfbam_cigar_expand(cigar.data(), (int32_t)cigar.size(), pos, introns ? 1 : 0, out.data(), (int32_t)cigar.size() + 1, &n,
                                device, err, sizeof err))
            value_error(err);
        py::list l;
        for (int i = 0; i < n; i++) l.append(py::make_tuple(out[2 * i], out[2 * i + 1]));
        return l;
    }, py::arg("cigar"), py::arg("pos"), py::arg("introns") = false, py::arg("device") = 0);

    py::class_<PyReader>(m, "Reader", "open BAM + BAI once (reference src/bam_ext.rs:6 open_bam), run several jobs")
        .def(py::init<const std::string&, py::object>(), py::arg("filename"), py::arg("index_filename") = py::none())
        .def("set_device", &PyReader::set_device)
        .def("set_devices", &PyReader::set_devices, "CUDA ordinals one call fans out over (default: see default_devices)")
        .def("devices", &PyReader::get_devices)
        .def("set_shard", &PyReader::set_shard, py::arg("index"), py::arg("count"),
             "this reader counts only shard `index` of `count` chunk-range shards (one-process-per-GPU jobs)")
        .def("preload", &PyReader::preload)
        .def("drop_preload", &PyReader::drop_preload)
        .def("pin_file", &PyReader::pin_file, "page-lock a mapping of the file; False if the platform refuses")
        .def("unpin_file", &PyReader::unpin_file)
        .def("targets", &PyReader::targets)
        .def("chunks", &PyReader::chunks, py::arg("gene_intervals") = py::none())
        .def("plan", &PyReader::plan, py::arg("gene_intervals"), py::arg("shard_index") = 0, py::arg("shard_count") = 1)
        .def("count_reads", &PyReader::count_reads, py::arg("mode"), py::arg("intervals"), py::arg("gene_intervals"),
             py::arg("each_read_counts_once") = py::none())
        .def("count_reads_raw", &PyReader::count_reads_raw, py::arg("mode"), py::arg("intervals"), py::arg("gene_intervals"),
             py::arg("each_read_counts_once") = py::none())
        .def("assemble", &PyReader::assemble)
        .def("count_introns", &PyReader::count_introns)
        .def("quantify_gene_reads", &PyReader::quantify_gene_reads, py::arg("intervals"), py::arg("gene_intervals"))
        .def("calculate_duplicate_distribution", &PyReader::duplicate_distribution)
        .def("count_reads_primary_only_right_strand_only_by_barcode", &PyReader::count_by_barcode, py::arg("intervals"),
             py::arg("gene_intervals"), py::arg("umi_strategy"))
        .def("stats", &PyReader::stats);
}
