// ref_driver.cpp -- C entry points around the UNMODIFIED reference objective.
//
// TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/tmb_shim/TMB.hpp).  This translation unit
// #includes /root/reference/src/smoothSDE.cpp where it lies (nothing of the reference is copied
// into this repo): `-I oracle/tmb_shim` resolves its `#include <TMB.hpp>` to the shim,
// `-I /root/reference/src` finds smoothSDE.cpp, whose own quoted includes pull in
// src/nllk/{nllk_sde,tr_dens,nllk_bm_ssm,nllk_ou_ssm,nllk_ctcrw,nllk_e_seal_ssm}.hpp.
// `objective_function<Type>::operator()` (smoothSDE.cpp:9-28) is instantiated three times:
//   Type = double                 value, REPORT(aest_all)
//   Type = ad::Var<double>        value + gradient (one reverse sweep of the tape)
//   Type = ad::Var<ad::Dual>      gradient + Hessian-vector product (reverse over forward)
// Built by oracle/Makefile into oracle/_ref/libsmoothsde_ref.so; Python side: oracle/oracle_ref.py.
#include <TMB.hpp>
#include "smoothSDE.cpp"

extern "C" {

// one named entry of the data list that SDE$setup() hands to MakeADFun (R/sde.R:528-598)
struct ssde_ref_item {
    const char* name;
    int kind;               // 0 double array, 1 int array, 2 string, 3 sparse triplets (0-based i, j, x)
    int ndim;
    long dim[3];
    const double* d;
    const int* i;
    const char* s;
    const int* ti;
    const int* tj;
    long nnz;
};
// parameter list in PARAMETER order, e.g. log_sigma_obs(1), coeff_fe, log_lambda, coeff_re
struct ssde_ref_par {
    const char* name;
    long len;
};
}

namespace {
std::map<std::string, shim_item> make_table(const ssde_ref_item* items, int n_items) {
    std::map<std::string, shim_item> t;
    for (int k = 0; k < n_items; k++) {
        shim_item it;
        it.kind = items[k].kind;
        for (int a = 0; a < items[k].ndim; a++) it.dim.push_back(items[k].dim[a]);
        it.d = items[k].d;
        it.i = items[k].i;
        if (items[k].s) it.s = items[k].s;
        it.ti = items[k].ti;
        it.tj = items[k].tj;
        it.nnz = items[k].nnz;
        t[items[k].name] = it;
    }
    return t;
}
void put_err(char* err, int errlen, const char* msg) {
    if (err && errlen > 0) {
        std::strncpy(err, msg, (size_t)errlen - 1);
        err[errlen - 1] = 0;
    }
}
}  // namespace

extern "C" {

// order 0: value.  order 1: value + gradient.  order 2: value + gradient + H * dir.
// Returns 0, or 1 with a message in err (R's error() / a shape problem).
int ssde_ref_eval(const ssde_ref_item* items, int n_items, const ssde_ref_par* layout, int n_layout,
                  const double* par, int order, const double* dir, double* value, double* grad, double* hvp,
                  char* err, int errlen) {
    try {
        auto table = make_table(items, n_items);
        if (order == 0) {
            objective_function<double> f;
            f.data = &table;
            long o = 0;
            for (int k = 0; k < n_layout; k++) {
                f.par[layout[k].name] = std::vector<double>(par + o, par + o + layout[k].len);
                o += layout[k].len;
            }
            *value = f();
        } else if (order == 1) {
            typedef ad::Var<double> V;
            auto& tape = ad::Tape<double>::get();
            tape.reset();
            objective_function<V> f;
            f.data = &table;
            std::vector<uint32_t> idx;
            long o = 0;
            for (int k = 0; k < n_layout; k++) {
                std::vector<V> p;
                for (long j = 0; j < layout[k].len; j++) { p.push_back(V::independent(par[o + j])); idx.push_back(p.back().i); }
                f.par[layout[k].name] = p;
                o += layout[k].len;
            }
            V y = f();
            *value = y.v;
            std::vector<double> adj;
            ad::reverse(y, adj);
            for (size_t j = 0; j < idx.size(); j++) grad[j] = adj[idx[j]];
            tape.reset();
        } else {
            typedef ad::Var<ad::Dual> V;
            auto& tape = ad::Tape<ad::Dual>::get();
            tape.reset();
            objective_function<V> f;
            f.data = &table;
            std::vector<uint32_t> idx;
            long o = 0;
            for (int k = 0; k < n_layout; k++) {
                std::vector<V> p;
                for (long j = 0; j < layout[k].len; j++) { p.push_back(V::independent(ad::Dual(par[o + j], dir[o + j]))); idx.push_back(p.back().i); }
                f.par[layout[k].name] = p;
                o += layout[k].len;
            }
            V y = f();
            *value = y.v.v;
            std::vector<ad::Dual> adj;
            ad::reverse(y, adj);
            for (size_t j = 0; j < idx.size(); j++) { if (grad) grad[j] = adj[idx[j]].v; hvp[j] = adj[idx[j]].d; }
            tape.reset();
        }
        return 0;
    } catch (const std::exception& e) {
        put_err(err, errlen, e.what());
        return 1;
    } catch (...) {
        put_err(err, errlen, "unknown C++ exception");
        return 1;
    }
}

// REPORT()ed object `name` of the double evaluation (aest_all, nllk_ctcrw.hpp:249), column-major
int ssde_ref_report(const ssde_ref_item* items, int n_items, const ssde_ref_par* layout, int n_layout,
                    const double* par, const char* name, double* out, long out_len, char* err, int errlen) {
    try {
        auto table = make_table(items, n_items);
        std::map<std::string, shim_report> rep;
        objective_function<double> f;
        f.data = &table;
        f.reports = &rep;
        long o = 0;
        for (int k = 0; k < n_layout; k++) {
            f.par[layout[k].name] = std::vector<double>(par + o, par + o + layout[k].len);
            o += layout[k].len;
        }
        f();
        auto it = rep.find(name);
        if (it == rep.end()) { put_err(err, errlen, "nothing REPORTed under that name"); return 1; }
        if ((long)it->second.x.size() != out_len) { put_err(err, errlen, "REPORT size mismatch"); return 1; }
        std::memcpy(out, it->second.x.data(), sizeof(double) * (size_t)out_len);
        return 0;
    } catch (const std::exception& e) {
        put_err(err, errlen, e.what());
        return 1;
    } catch (...) {
        put_err(err, errlen, "unknown C++ exception");
        return 1;
    }
}

// size of one tape node, for memory planning on the Python side
int ssde_ref_tape_node_bytes(void) { return (int)sizeof(ad::Tape<double>::Node); }
}
