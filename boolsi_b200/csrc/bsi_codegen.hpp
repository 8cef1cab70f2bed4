// Network / problem-space analysis and the generator of the per-network macro header that
// specialises the hand-written kernel template (bsi_template.cuh) through NVRTC.
//
// What it replaces in the reference: the truth-table dict lookup of model.py:26-28 becomes, per
// node, a chain of lop3.b32 instructions on bit-sliced words (32 trajectories per word, one
// register word per node); adjust_update_rules_for_fixed_nodes (model.py:31-49) and the
// perturbation override (model.py:70-71) become constants / AND-OR masks; the mixed-radix problem
// enumeration (batching.py:160-229) becomes index arithmetic emitted here.
#pragma once

#include <algorithm>
#include <cstdint>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/bsi.h"

namespace bsi {

struct BoolFn {
    std::vector<int> vars;        // node ids, vars[0] is the most significant table index bit
    std::vector<uint8_t> tt;      // 2^k entries

    int k() const { return (int)vars.size(); }
    std::string key() const {
        std::string s;
        for (int v : vars) { s += std::to_string(v); s += ','; }
        s += ':';
        for (uint8_t b : tt) s += b ? '1' : '0';
        return s;
    }
    BoolFn cofactor(int m, int val) const {
        BoolFn g;
        int kk = k();
        for (int i = 0; i < kk; ++i) if (i != m) g.vars.push_back(vars[i]);
        int shift = kk - 1 - m;
        g.tt.resize((size_t)1 << (kk - 1));
        for (size_t j = 0; j < g.tt.size(); ++j) {
            size_t low = j & (((size_t)1 << shift) - 1);
            size_t high = j >> shift;
            size_t full = (high << (shift + 1)) | ((size_t)val << shift) | low;
            g.tt[j] = tt[full];
        }
        return g;
    }
    bool depends_on(int m) const { return cofactor(m, 0).tt != cofactor(m, 1).tt; }
    BoolFn reduced() const {
        BoolFn f = *this;
        for (int m = 0; m < f.k();) {
            if (!f.depends_on(m)) f = f.cofactor(m, 0); else ++m;
        }
        return f;
    }
};

struct Operand {
    enum Kind { CONST, REG } kind = CONST;
    int value = 0;            // CONST
    std::string reg;          // REG: expression text (e.g. "S##12" or "t7_")
    bool neg = false;
};

// Emits the LOP3 chain of one UPDATE macro body.
class Lop3Synth {
  public:
    explicit Lop3Synth(const std::string &src_prefix) : src_(src_prefix) {}

    // returns the text of statements computing the function; result operand in `out`
    Operand synth(const BoolFn &f_in) {
        BoolFn f = f_in.reduced();
        if (f.k() == 0) { Operand o; o.kind = Operand::CONST; o.value = f.tt[0]; return o; }
        if (f.k() == 1) {
            Operand o; o.kind = Operand::REG; o.reg = src_ + "##" + std::to_string(f.vars[0]);
            o.neg = (f.tt[0] == 1);
            return o;
        }
        std::string key = f.key();
        auto it = memo_.find(key);
        if (it != memo_.end()) return it->second;
        Operand res;
        res.kind = Operand::REG;
        res.reg = "t" + std::to_string(n_tmp_++) + "_";
        if (f.k() <= 3) {
            std::string r[3];
            for (int i = 0; i < 3; ++i) r[i] = src_ + "##" + std::to_string(f.vars[std::min(i, f.k() - 1)]);
            unsigned lut = 0;
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c) {
                int idx = f.k() == 3 ? (a << 2 | b << 1 | c) : (a << 1 | b);
                if (f.tt[idx]) lut |= 1u << (a << 2 | b << 1 | c);
            }
            emit(res.reg, lut, r[0], r[1], r[2]);
        } else {
            int m = best_split(f);
            Operand hi = synth(f.cofactor(m, 1));
            Operand lo = synth(f.cofactor(m, 0));
            std::string sel = src_ + "##" + std::to_string(f.vars[m]);
            std::string rb = hi.kind == Operand::REG ? hi.reg : sel;
            std::string rc = lo.kind == Operand::REG ? lo.reg : sel;
            unsigned lut = 0;
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c) {
                int vb = hi.kind == Operand::CONST ? hi.value : (b ^ (hi.neg ? 1 : 0));
                int vc = lo.kind == Operand::CONST ? lo.value : (c ^ (lo.neg ? 1 : 0));
                if (a ? vb : vc) lut |= 1u << (a << 2 | b << 1 | c);
            }
            emit(res.reg, lut, sel, rb, rc);
        }
        memo_[key] = res;
        return res;
    }

    // statement assigning the operand to `dst`
    void assign(const std::string &dst, const Operand &o) {
        if (o.kind == Operand::CONST) {
            body_ << dst << " = " << (o.value ? "0xFFFFFFFFu" : "0u") << "; ";
        } else if (o.neg) {
            body_ << dst << " = ~" << o.reg << "; ";
            ++n_lop3_;
        } else {
            body_ << dst << " = " << o.reg << "; ";
        }
    }
    // dst = (value & ~fm) | fv   -- per-lane fixed-node masks
    void assign_masked(const std::string &dst, const Operand &o, const std::string &fm,
                       const std::string &fv) {
        if (o.kind == Operand::CONST) {
            // (c & ~fm) | fv
            body_ << dst << " = " << (o.value ? "(~" + fm + ") | " + fv : fv) << "; ";
            n_lop3_ += o.value ? 1 : 0;
        } else {
            unsigned lut = o.neg ? 0xABu /* (~a & ~b) | c */ : 0xBAu /* (a & ~b) | c */;
            body_ << dst << " = bsi_lop3<0x" << std::hex << lut << std::dec << ">(" << o.reg << ", "
                  << fm << ", " << fv << "); ";
            ++n_lop3_;
        }
    }
    std::string body() const { return body_.str(); }
    std::string temp_decls() const {
        std::ostringstream o;
        for (int i = 0; i < n_tmp_; ++i) o << (i ? ", " : "unsigned int ") << "t" << i << "_";
        if (n_tmp_) o << "; ";
        return o.str();
    }
    int n_lop3() const { return n_lop3_; }

  private:
    void emit(const std::string &dst, unsigned lut, const std::string &a, const std::string &b,
              const std::string &c) {
        body_ << dst << " = bsi_lop3<0x" << std::hex << lut << std::dec << ">(" << a << ", " << b
              << ", " << c << "); ";
        ++n_lop3_;
    }
    int cost(const BoolFn &f_in) {
        BoolFn f = f_in.reduced();
        if (f.k() <= 1) return 0;
        if (f.k() <= 3) return 1;
        std::string key = f.key();
        auto it = cost_memo_.find(key);
        if (it != cost_memo_.end()) return it->second;
        int best = 1 << 30;
        if (f.k() > 8) {
            best = cost(f.cofactor(0, 0)) + cost(f.cofactor(0, 1)) + 1;
        } else {
            for (int m = 0; m < f.k(); ++m)
                best = std::min(best, cost(f.cofactor(m, 0)) + cost(f.cofactor(m, 1)) + 1);
        }
        cost_memo_[key] = best;
        return best;
    }
    int best_split(const BoolFn &f) {
        if (f.k() > 8) return 0;
        int best = 1 << 30, arg = 0;
        for (int m = 0; m < f.k(); ++m) {
            int c = cost(f.cofactor(m, 0)) + cost(f.cofactor(m, 1)) + 1;
            if (c < best) { best = c; arg = m; }
        }
        return arg;
    }

    std::string src_;
    std::ostringstream body_;
    std::map<std::string, Operand> memo_;
    std::map<std::string, int> cost_memo_;
    int n_tmp_ = 0;
    int n_lop3_ = 0;
};

struct Variation {
    int kind;        // bsi_variation_kind
    int node;
    uint64_t t;
    int range;       // bsi_range
    int radix() const { return range == BSI_RANGE_MAYBE_TRUE_OR_FALSE ? 3 : 2; }
};

struct Perturbation { uint64_t t; int node; int value; };

// Host copy of the descriptors, validated.
struct Model {
    int n = 0;
    int kw = 1;
    std::vector<BoolFn> fn;                 // per node
    std::vector<uint8_t> initial;           // origin initial state
    std::map<int, int> fixed;               // origin fixed nodes
    std::vector<Perturbation> perts;        // origin perturbations
    std::vector<Variation> vars;
    int n_init_vars = 0;                    // leading BSI_VAR_INITIAL variations
    bool lanes_uniform = true;              // no fixed/perturbation variations
    uint64_t n_problems = 1;
    bool n_problems_overflow = false;
    uint64_t lp_max = 0;                    // max perturbation time over origin + variations
    uint64_t lp_origin = 0;                 // max perturbation time of the origin alone
    int mode = BSI_MODE_ATTRACT;
    bsi_target_desc target{};
    std::vector<int> target_nodes;
};

inline std::string load_model(const bsi_network_desc *net, const bsi_space_desc *sp, int mode,
                              const bsi_target_desc *target, Model &m) {
    if (!net || !sp) return "null descriptor";
    if (net->n_nodes == 0) return "network has no nodes";
    if (net->n_nodes > 64 * BSI_KEY_WORDS_MAX) return "n_nodes > 256 is not supported";
    if (!net->pred_off || !net->tt_off || !net->tt || !sp->initial_state) return "null array in descriptor";
    m.n = (int)net->n_nodes;
    m.kw = (m.n + 63) / 64;
    m.mode = mode;
    m.fn.resize(m.n);
    for (int i = 0; i < m.n; ++i) {
        uint32_t b = net->pred_off[i], e = net->pred_off[i + 1];
        if (e < b || e - b > 16) return "a node has more than 16 inputs (unsupported)";
        BoolFn &f = m.fn[i];
        for (uint32_t q = b; q < e; ++q) {
            if (net->pred_idx[q] >= net->n_nodes) return "predecessor index out of range";
            f.vars.push_back((int)net->pred_idx[q]);
        }
        uint64_t tb = net->tt_off[i], te = net->tt_off[i + 1];
        if (te - tb != ((uint64_t)1 << f.vars.size())) return "truth table size does not match 2^K";
        f.tt.assign(net->tt + tb, net->tt + te);
        for (auto &x : f.tt) x = x ? 1 : 0;
        // duplicate predecessors are legal in principle: fold them (keep first occurrence)
        for (int a = 0; a < f.k(); ++a)
            for (int c = a + 1; c < f.k();) {
                if (f.vars[c] == f.vars[a]) {
                    // restrict to entries where both agree
                    BoolFn g;
                    g.vars = f.vars;
                    g.vars.erase(g.vars.begin() + c);
                    int kk = f.k();
                    g.tt.resize((size_t)1 << (kk - 1));
                    for (size_t j = 0; j < g.tt.size(); ++j) {
                        // build full index: insert at position c the value of var a
                        int sa_g = (kk - 2) - a;               // bit of var a in g's index
                        int va = (int)((j >> sa_g) & 1);
                        int shift = kk - 1 - c;
                        size_t low = j & (((size_t)1 << shift) - 1);
                        size_t high = j >> shift;
                        g.tt[j] = f.tt[(high << (shift + 1)) | ((size_t)va << shift) | low];
                    }
                    f = g;
                } else {
                    ++c;
                }
            }
    }
    m.initial.assign(sp->initial_state, sp->initial_state + m.n);
    for (auto &x : m.initial) x = x ? 1 : 0;
    for (uint32_t q = 0; q < sp->n_fixed; ++q) {
        if (sp->fixed_node[q] >= net->n_nodes) return "fixed node out of range";
        m.fixed[(int)sp->fixed_node[q]] = sp->fixed_value[q] ? 1 : 0;
    }
    for (uint32_t q = 0; q < sp->n_perturbations; ++q) {
        if (sp->pert_node[q] >= net->n_nodes) return "perturbed node out of range";
        if (sp->pert_t[q] == 0) return "perturbation at t = 0";
        m.perts.push_back({sp->pert_t[q], (int)sp->pert_node[q], sp->pert_value[q] ? 1 : 0});
        m.lp_origin = std::max(m.lp_origin, sp->pert_t[q]);
    }
    m.lp_max = m.lp_origin;
    bool seen_other = false;
    for (uint32_t q = 0; q < sp->n_variations; ++q) {
        Variation v{(int)sp->var_kind[q], (int)sp->var_node[q], sp->var_t[q], (int)sp->var_range[q]};
        if (v.node < 0 || v.node >= m.n) return "variation node out of range";
        if (v.kind < 0 || v.kind > 2 || v.range < 0 || v.range > 3) return "bad variation";
        if (v.kind == BSI_VAR_INITIAL) {
            if (seen_other) return "initial-state variations must come first";
            if (v.range != BSI_RANGE_TRUE_OR_FALSE) return "initial-state variation must be TRUE_OR_FALSE";
            m.n_init_vars++;
        } else {
            seen_other = true;
            m.lanes_uniform = false;
            if (v.kind == BSI_VAR_PERTURBATION) {
                if (v.t == 0) return "perturbation variation at t = 0";
                m.lp_max = std::max(m.lp_max, v.t);
            }
        }
        uint64_t r = (uint64_t)v.radix();
        if (m.n_problems > UINT64_MAX / r) m.n_problems_overflow = true; else m.n_problems *= r;
        m.vars.push_back(v);
    }
    if (m.n_problems_overflow) return "problem space exceeds 2^64";
    if (mode == BSI_MODE_ATTRACT && !m.lanes_uniform)
        return "attract: fixed-node / perturbation variations are not allowed (reference input.py:392-397, 537-543)";
    if (mode == BSI_MODE_TARGET) {
        if (!target) return "target mode needs a target descriptor";
        m.target = *target;
        for (int i = 0; i < m.n; ++i)
            if ((target->mask[i >> 6] >> (i & 63)) & 1) m.target_nodes.push_back(i);
    }
    return "";
}

struct GenInfo {
    int filter2_nodes = 0;   // stage-2 filter words per checkpoint slot (shared memory)
    int update_lop3 = 0;     // LOP3 per word-step of the autonomous update
    int update_quiet_lop3 = 0;   // the same past the last perturbation time (fixed nodes folded into their readers)
    int compare_lop3 = 0;    // full compare
    int filter_lop3 = 0;     // cheap pre-compare
    std::vector<int> filter_nodes;
};

inline std::string node_list_macro(const std::string &name, int n) {
    std::ostringstream o;
    o << "#define " << name << "(M)";
    for (int i = 0; i < n; ++i) o << " M(" << i << ")";
    o << "\n";
    return o.str();
}

// The macro header.  `filter` = nodes of the cheap pre-compare, in order.
inline std::string generate_header(const Model &m, const std::vector<int> &filter,
                                   const std::vector<int> &filter2, int block_threads,
                                   int min_blocks, int debug_stats, GenInfo &info) {
    std::ostringstream o;
    const int n = m.n;
    o << "// generated by libbsi for a " << n << "-node network\n";
    o << "#define BSI_N " << n << "\n#define BSI_KW " << m.kw << "\n";
    o << "#define BSI_BLOCK " << block_threads << "\n";
    o << "#define BSI_MIN_BLOCKS " << min_blocks << "\n";
    o << "#define BSI_STATS " << (debug_stats ? 1 : 0) << "\n";
    o << "#define BSI_MODE " << m.mode << "\n";
    o << "#define BSI_LANES_UNIFORM " << (m.lanes_uniform ? 1 : 0) << "\n";
    o << "#define BSI_LP_ORIGIN " << m.lp_origin << "ull\n";
    o << "#define BSI_LP_MAX " << m.lp_max << "ull\n";
    o << node_list_macro("BSI_FOR_NODES", n);
    {   // reverse order (most significant node first) for the lexicographic minimum
        o << "#define BSI_FOR_NODES_DESC(M)";
        for (int i = n - 1; i >= 0; --i) o << " M(" << i << ")";
        o << "\n";
    }

    // nodes whose fixing varies per lane
    std::set<int> fixvar_nodes;
    for (const auto &v : m.vars) if (v.kind == BSI_VAR_FIXED) fixvar_nodes.insert(v.node);
    o << "#define BSI_FOR_FIXVAR(M)";
    for (int node : fixvar_nodes) o << " M(" << node << ")";
    o << "\n";
    // perturbation variations get one (mask, value) word pair each
    int n_pvar = 0;
    for (const auto &v : m.vars) if (v.kind == BSI_VAR_PERTURBATION) ++n_pvar;
    o << "#define BSI_FOR_PVAR(M)";
    for (int q = 0; q < n_pvar; ++q) o << " M(" << q << ")";
    o << "\n";

    // ---- update: D = F(S) with fixed nodes folded in (model.py:16-49) ---------------------
    {
        Lop3Synth syn("S");
        // Emission order: greedily pick the node whose evaluation retires the most old-state words
        // (a source dies once its last reader has been emitted), so that the old and the new
        // state overlap in as few registers as possible.
        std::vector<int> order;
        {
            std::vector<std::vector<int>> srcs(n);
            std::vector<int> readers(n, 0);
            for (int i = 0; i < n; ++i) {
                bool constant = m.fixed.count(i) && !fixvar_nodes.count(i);
                if (!constant) {
                    BoolFn r = m.fn[i].reduced();
                    std::set<int> uniq(r.vars.begin(), r.vars.end());
                    srcs[i].assign(uniq.begin(), uniq.end());
                }
                for (int v : srcs[i]) ++readers[v];
            }
            std::vector<char> done(n, 0);
            for (int step = 0; step < n; ++step) {
                int best = -1, best_score = -1000;
                for (int i = 0; i < n; ++i) {
                    if (done[i]) continue;
                    int kills = 0;
                    for (int v : srcs[i]) if (readers[v] == 1) ++kills;
                    int score = 4 * kills - (int)srcs[i].size();
                    if (score > best_score) { best_score = score; best = i; }
                }
                done[best] = 1;
                order.push_back(best);
                for (int v : srcs[best]) --readers[v];
            }
        }
        for (int i : order) {
            std::string dst = "D##" + std::to_string(i);
            auto fx = m.fixed.find(i);
            bool varfix = fixvar_nodes.count(i) > 0;
            if (fx != m.fixed.end() && !varfix) {
                Operand c; c.kind = Operand::CONST; c.value = fx->second;
                syn.assign(dst, c);
            } else if (varfix) {
                // origin fixing (if any) is folded into the lane masks by the prologue
                Operand v = syn.synth(m.fn[i]);
                syn.assign_masked(dst, v, "fm" + std::to_string(i), "fv" + std::to_string(i));
            } else {
                Operand v = syn.synth(m.fn[i]);
                syn.assign(dst, v);
            }
        }
        o << "#define BSI_UPDATE(D, S) { " << syn.temp_decls() << syn.body() << "}\n";
        info.update_lop3 = syn.n_lop3();

        // The same update for a state that is known to carry its fixed values: a node fixed for all lanes
        // holds its constant from time 1 on unless a perturbation has just overwritten it (model.py:31-49,
        // 66-71), so past the last perturbation time the readers of such a node can take the constant
        // instead of the register -- one input fewer per reader, and the node's word leaves the loop.
        Lop3Synth quiet("S");
        for (int i : order) {
            std::string dst = "D##" + std::to_string(i);
            auto fx = m.fixed.find(i);
            bool varfix = fixvar_nodes.count(i) > 0;
            if (fx != m.fixed.end() && !varfix) {
                Operand c; c.kind = Operand::CONST; c.value = fx->second;
                quiet.assign(dst, c);
                continue;
            }
            BoolFn f = m.fn[i];
            for (int q = 0; q < f.k();) {
                auto src = m.fixed.find(f.vars[q]);
                if (src != m.fixed.end() && !fixvar_nodes.count(f.vars[q])) f = f.cofactor(q, src->second);
                else ++q;
            }
            Operand v = quiet.synth(f);
            if (varfix) quiet.assign_masked(dst, v, "fm" + std::to_string(i), "fv" + std::to_string(i));
            else quiet.assign(dst, v);
        }
        o << "#define BSI_UPDATE_QUIET(D, S) { " << quiet.temp_decls() << quiet.body() << "}\n";
        info.update_quiet_lop3 = quiet.n_lop3();
    }

    // ---- perturbations scheduled at time t (model.py:70-71), uniform part + per-lane part ----
    {
        std::map<uint64_t, std::string> by_t;
        for (const auto &p : m.perts) {
            std::ostringstream s;
            s << "X##" << p.node << " = " << (p.value ? "0xFFFFFFFFu" : "0u") << "; ";
            by_t[p.t] += s.str();
        }
        int q = 0;
        for (const auto &v : m.vars) {
            if (v.kind != BSI_VAR_PERTURBATION) continue;
            std::ostringstream s;
            s << "X##" << v.node << " = bsi_lop3<0xBA>(X##" << v.node << ", pm" << q << ", pv" << q << "); ";
            by_t[v.t] += s.str();
            ++q;
        }
        o << "#define BSI_PERTURB(X, t) switch (t) { ";
        for (auto &kv : by_t) o << "case " << kv.first << "ull: " << kv.second << "break; ";
        o << "default: break; }\n";
        // per-lane "all perturbations carried out" masks for target mode (model.py:125, 178-187)
        o << "#define BSI_LP_TIMES(M)";
        std::set<uint64_t> times;
        for (auto &kv : by_t) times.insert(kv.first);
        for (uint64_t t : times) o << " M(" << t << "ull)";
        o << "\n";
    }

    // ---- compare: acc |= X ^ Y over the filter nodes (two accumulators to halve the dependency
    //      chain), then over the rest -- against registers or against the checkpoint scratch -----
    {
        std::vector<char> in_filter(n, 0);
        for (int f : filter) in_filter[f] = 1;
        o << "#define BSI_FOR_FILTER(M)";
        for (int f : filter) o << " M(" << f << ")";
        o << "\n#define BSI_FOR_REST(M)";
        for (int i = 0; i < n; ++i) if (!in_filter[i]) o << " M(" << i << ")";
        o << "\n#define BSI_CMP_FILTER(acc, acc2, X, Y) {";
        int parity = 0;
        for (int f : filter) {
            const char *acc = (parity++ & 1) ? "acc2" : "acc";
            o << " " << acc << " = bsi_lop3<0xF6>(" << acc << ", X##" << f << ", Y##" << f << ");";
        }
        // the hot-loop form: acc = ~ALIVE | OR_f (X_f ^ Y_f) as ONE chain of |filter| lop3 -- the
        // first one folds the "not alive" term in (0x6F = ~a | (b ^ c)).  The chain is short next
        // to the ~100 independent lop3 of a step, so a second accumulator (and the two extra
        // instructions to start and merge it) buys nothing.
        o << " }\n#define BSI_CMP_FILTER1(acc, ALIVE, X, Y) {";
        if (filter.empty()) o << " acc = ~(ALIVE);";
        for (size_t g = 0; g < filter.size(); ++g) {
            if (g == 0) o << " acc = bsi_lop3<0x6F>(ALIVE, X##" << filter[g] << ", Y##" << filter[g] << ");";
            else o << " acc = bsi_lop3<0xF6>(acc, X##" << filter[g] << ", Y##" << filter[g] << ");";
        }
        o << " }\n#define BSI_CMP_REST(acc, X, Y) {";
        for (int i = 0; i < n; ++i)
            if (!in_filter[i]) o << " acc = bsi_lop3<0xF6>(acc, X##" << i << ", Y##" << i << ");";
        // stage 2 in two tiers: a register-filter pass is almost always ONE lane that differs from
        // the checkpoint elsewhere, and four more words rule it out 15 times out of 16
        const size_t tier_a = std::min<size_t>(4, filter2.size());
        o << " }\n#define BSI_CMP_FILTER2A_MEM(acc, X, P, S) {";
        for (size_t g = 0; g < tier_a; ++g)
            o << " acc = bsi_lop3<0xF6>(acc, X##" << filter2[g] << ", (P)[" << filter2[g] << " * (S)]);";
        o << " }\n#define BSI_CMP_FILTER2B_MEM(acc, X, P, S) {";
        for (size_t g = tier_a; g < filter2.size(); ++g)
            o << " acc = bsi_lop3<0xF6>(acc, X##" << filter2[g] << ", (P)[" << filter2[g] << " * (S)]);";
        // full compare of a register state against a checkpoint in memory: one load + one lop3 per node
        o << " }\n#define BSI_CMP_ALL_MEM(acc, X, P, S) {";
        for (int i = 0; i < n; ++i)
            o << " acc = bsi_lop3<0xF6>(acc, X##" << i << ", (P)[" << i << " * (S)]);";
        o << " }\n#define BSI_STORE_ALL(X, P, S) {";
        for (int i = 0; i < n; ++i) o << " (P)[" << i << " * (S)] = X##" << i << ";";
        o << " }\n#define BSI_LOAD_MEM(X, P, S) {";
        for (int i = 0; i < n; ++i) o << " X##" << i << " = (P)[" << i << " * (S)];";
        o << " }\n#define BSI_LOAD_STREAM(X, P, S) {";
        for (int i = 0; i < n; ++i) o << " X##" << i << " = bsi_ld_stream((P) + " << i << " * (S));";
        o << " }\n#define BSI_DUMP(X, A) {";
        for (int i = 0; i < n; ++i) o << " (A)[" << i << "] = X##" << i << ";";
        o << " }\n";
        // lexicographic compare, most significant node first, in chunks of 8 with early exit
        o << "#define BSI_LEX_CHUNKS";
        int open = 0;
        for (int i = n - 1; i >= 0; --i) {
            o << " BSI_LEX_NODE(" << i << ")";
            if (i > 0 && (n - i) % 8 == 0) { o << " if (eq_) {"; ++open; }
        }
        for (int q = 0; q < open; ++q) o << " }";
        o << "\n";
        info.filter_lop3 = (int)filter.size();
        info.filter2_nodes = (int)std::max<size_t>(1, filter2.size());
        info.compare_lop3 = n;
        info.filter_nodes = filter;
    }
    // target compare: acc |= (X ^ code) on the constrained nodes
    if (m.mode == BSI_MODE_TARGET) {
        o << "#define BSI_CMP_TARGET(acc, X) {";
        // acc |= X ^ code bit: two constrained nodes per lop3, acc | (b ^ cb) | (c ^ cc)
        auto want = [&](int node) { return (int)((m.target.code[node >> 6] >> (node & 63)) & 1); };
        size_t q = 0;
        for (; q + 1 < m.target_nodes.size(); q += 2) {
            int nb = m.target_nodes[q], nc = m.target_nodes[q + 1];
            unsigned lut = 0;
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int c = 0; c < 2; ++c)
                if (a | (b ^ want(nb)) | (c ^ want(nc))) lut |= 1u << (a << 2 | b << 1 | c);
            o << " acc = bsi_lop3<0x" << std::hex << lut << std::dec << ">(acc, X##" << nb << ", X##" << nc << ");";
        }
        if (q < m.target_nodes.size()) {
            int node = m.target_nodes[q];
            o << " acc |= " << (want(node) ? "~" : "") << "X##" << node << ";";
        }
        o << " }\n";
    }

    // ---- initial state -------------------------------------------------------------------------
    // origin code words (free nodes zeroed)
    std::vector<uint64_t> origin(m.kw, 0);
    for (int i = 0; i < n; ++i) if (m.initial[i]) origin[i >> 6] |= (uint64_t)1 << (i & 63);
    for (int q = 0; q < m.n_init_vars; ++q)
        origin[m.vars[q].node >> 6] &= ~((uint64_t)1 << (m.vars[q].node & 63));

    // dense: word W holds problems 32W .. 32W+31; valid only when all variations are binary
    // initial-state digits (always so in attract mode)
    {
        static const char *lane_pat[5] = {"0xAAAAAAAAu", "0xCCCCCCCCu", "0xF0F0F0F0u",
                                          "0xFF00FF00u", "0xFFFF0000u"};
        std::vector<std::string> expr(n);
        for (int i = 0; i < n; ++i) expr[i] = m.initial[i] ? "0xFFFFFFFFu" : "0u";
        for (int q = 0; q < m.n_init_vars; ++q) {
            int node = m.vars[q].node;
            if (q < 5) expr[node] = lane_pat[q];
            else expr[node] = "(0u - (unsigned int)(((W) >> " + std::to_string(q - 5) + ") & 1ull))";
        }
        o << "#define BSI_INIT_DENSE(X, W) {";
        for (int i = 0; i < n; ++i) o << " X##" << i << " = " << expr[i] << ";";
        o << " }\n";
    }
    // per lane: problem index -> initial code words c[0..KW) and lane-mask bits
    {
        o << "#define BSI_LANE_SETUP(l, idx, c) { unsigned long long r_ = (idx); unsigned int d_; ";
        for (int w = 0; w < m.kw; ++w) o << "c[" << w << "] = 0x" << std::hex << origin[w] << std::dec << "ull; ";
        int pq = 0;
        // origin fixing of FIXVAR nodes goes into the masks first
        for (int node : fixvar_nodes) {
            auto fx = m.fixed.find(node);
            if (fx != m.fixed.end()) {
                o << "fm" << node << " |= 1u << (l); ";
                if (fx->second) o << "fv" << node << " |= 1u << (l); ";
            }
        }
        for (const auto &v : m.vars) {
            if (v.radix() == 2) o << "d_ = (unsigned int)(r_ & 1ull); r_ >>= 1; ";
            else o << "d_ = (unsigned int)(r_ % 3ull); r_ /= 3ull; ";
            if (v.kind == BSI_VAR_INITIAL) {
                o << "c[" << (v.node >> 6) << "] |= (unsigned long long)d_ << " << (v.node & 63) << "; ";
                continue;
            }
            // on = variation active for this lane; val = its value
            std::string on, val;
            switch (v.range) {
            case BSI_RANGE_TRUE_OR_FALSE: on = "1u"; val = "d_"; break;
            case BSI_RANGE_MAYBE_FALSE: on = "d_"; val = "0u"; break;
            case BSI_RANGE_MAYBE_TRUE: on = "d_"; val = "d_"; break;
            default: on = "(d_ != 0u)"; val = "(d_ == 2u)"; break;
            }
            if (v.kind == BSI_VAR_FIXED) {
                // a later variation of the same node overrides an earlier one / the origin
                o << "if (" << on << ") { fm" << v.node << " |= 1u << (l); fv" << v.node
                  << " = (fv" << v.node << " & ~(1u << (l))) | ((unsigned int)" << val << " << (l)); } ";
            } else {
                o << "pm" << pq << " |= (unsigned int)" << on << " << (l); pv" << pq
                  << " |= (unsigned int)((" << on << ") & (" << val << ")) << (l); ";
                // per-lane last perturbation time
                o << "if (" << on << " && " << v.t << "ull > lp_lane_) lp_lane_ = " << v.t << "ull; ";
                ++pq;
            }
        }
        o << "}\n";
    }
    // code words <-> bit-sliced words: all 32 lanes at once through 32x32 bit-matrix transposes
    // (bsi_transpose32 maps B[r] bit c = A[31-c] bit 31-r, hence the reversed indices): CW is
    // u64 [32 lanes][KW].
    {
        const int blocks = (n + 31) / 32;
        o << "#define BSI_PACK(X, CW) { unsigned int A_[32];";
        for (int b = 0; b < blocks; ++b) {
            o << " _Pragma(\"unroll\") for (int k_ = 0; k_ < 32; ++k_) A_[k_] = (unsigned int)((CW)[31 - k_]["
              << (b >> 1) << "] >> " << (32 * (b & 1)) << "); BSI_TRANSPOSE32(A_);";
            for (int i = 0; i < 32 && 32 * b + i < n; ++i)
                o << " X##" << (32 * b + i) << " = A_[" << (31 - i) << "];";
        }
        o << " }\n#define BSI_UNPACK(CW, X) { unsigned int A_[32];";
        for (int b = 0; b < blocks; ++b) {
            for (int k = 0; k < 32; ++k) {
                int node = 32 * b + 31 - k;
                if (node < n) o << " A_[" << k << "] = X##" << node << ";";
                else o << " A_[" << k << "] = 0u;";
            }
            o << " BSI_TRANSPOSE32(A_); _Pragma(\"unroll\") for (int k_ = 0; k_ < 32; ++k_) (CW)[31 - k_]["
              << (b >> 1) << "] " << ((b & 1) ? "|= (unsigned long long)A_[k_] << 32;" : "= (unsigned long long)A_[k_];");
        }
        o << " }\n";
    }
    return o.str();
}

}  // namespace bsi
