// include/scpp_plugin.hpp — lowers the application constraints a model records through the cvx:: shim (include/scpp_cvx.hpp;
// reference: Model::addApplicationConstraints, scpp_core/include/systemModel.hpp:76-78, e.g. scpp_models/src/rocketQuat.cpp:70-144)
// to the STAGE-WISE tables the device kernels work on (scpp_b200/csrc/models.cuh):
//      * pinned variables        single-variable equalities  x_0 = x_init, final-state rows, u_{K-1,j} = 0, X.row(i) == 0
//      * LP rows                 s_r = h_r - sum_j c_j xi[idx_j] >= 0          (xi = [x ; u] of one node)
//      * second-order cones      (s_0 ; s_1..) in Q^d, every s_i of the same affine form
// with every coefficient traced back to where its value lives: a literal, an entry of the model's constant block (dynpar of a
// Parameters field), an entry of x_init / x_final, or a per-node parameter array (the linearised minimum-thrust direction).
// The engine's tables are per NODE and identical for all nodes, so the lowering checks that the recorded problem has that
// structure and reports what breaks it otherwise.  Host-only, header-only.
#pragma once
#include "scpp_cvx.hpp"
#include <algorithm>
#include <cstdio>
#include <sstream>

namespace scpp_plugin {

struct Region { const double *base; int n; };       // a block of doubles the model's dynpars may point into

struct Source {
    enum Kind { LITERAL, CONSTANT, X_INIT, X_FINAL, NODE_ARRAY } kind = LITERAL;
    double lit = 0.;      // LITERAL: the value
    int index = 0;        // CONSTANT / X_INIT / X_FINAL: entry ; NODE_ARRAY: row of the rows x K array (the node gives the column)
    double scale = 1.;
    double value(const double *constants, const double *x_init, const double *x_final, const double *node_array, int node_rows, int k) const
    {
        switch (kind) {
        case LITERAL: return lit;
        case CONSTANT: return scale * constants[index];
        case X_INIT: return scale * x_init[index];
        case X_FINAL: return scale * x_final[index];
        default: return scale * node_array[(size_t)k * node_rows + index];
        }
    }
    bool same(const Source &o) const { return kind == o.kind && index == o.index && scale == o.scale && (kind != LITERAL || lit == o.lit); }
};

struct Row {                                  // s = h - sum c_j xi[idx_j]
    std::vector<std::pair<int, Source>> c;
    Source h;
    bool same(const Row &o) const
    {
        if (c.size() != o.c.size() || !h.same(o.h)) return false;
        for (size_t i = 0; i < c.size(); i++) if (c[i].first != o.c[i].first || !c[i].second.same(o.c[i].second)) return false;
        return true;
    }
};
struct Pin { int idx; Source value; };

struct StageTable {
    int nx = 0, nu = 0, K = 0;
    std::vector<Row> lp;                      // LP rows of every node, in the order the model added them
    std::vector<std::vector<Row>> cones;      // cones of every node (head row first)
    std::vector<Pin> pin_all, pin_first, pin_last;
    int cone_rows() const { int n = 0; for (auto &c : cones) n += (int)c.size(); return n; }
    int max_cone_dim() const { int n = 0; for (auto &c : cones) n = std::max(n, (int)c.size()); return n; }
};

class Lowering {
public:
    Region constants{nullptr, 0}, x_init{nullptr, 0}, x_final{nullptr, 0}, node_array{nullptr, 0};
    int node_rows = 0;                        // node_array is node_rows x K, column-major (thrust_const, rocketQuat.cpp:115-116)

    StageTable lower(const cvx::OptimizationProblem &p, const std::string &xname = "X", const std::string &uname = "U")
    {
        const auto &vx = p.var(xname), &vu = p.var(uname);
        StageTable t;
        t.nx = vx.rows; t.nu = vu.rows; t.K = vx.cols;
        if (vu.cols != t.K) fail("first-order-hold layout expected: U has one column per node");
        X_ = vx; U_ = vu; K_ = t.K;
        std::vector<std::vector<Row>> lp(t.K);
        std::vector<std::vector<std::vector<Row>>> cones(t.K);
        std::vector<std::vector<Pin>> pins(t.K);
        for (const auto &c : p.constraints) {
            if (c.kind == cvx::Constraint::EQ) {
                int k; Row r = row_of(c.rhs - c.lhs, k);          // 0 = h - sum c xi
                if (r.c.size() != 1) fail("only single-variable equalities can be lowered (they pin a variable)");
                const Source &cf = r.c[0].second;
                if (cf.kind != Source::LITERAL || cf.lit == 0.) fail("pinned variable with a non-literal coefficient");
                Pin pin{r.c[0].first, r.h};
                if (pin.value.kind == Source::LITERAL) pin.value.lit /= cf.lit; else pin.value.scale /= cf.lit;
                pins[k].push_back(pin);
            } else if (c.kind == cvx::Constraint::LE) {
                int k; lp[k_of(c.rhs - c.lhs)].push_back(row_of(c.rhs - c.lhs, k));
            } else {
                std::vector<Row> cone;
                int k = -1, kk;
                cone.push_back(row_of(c.rhs, kk)); k = kk;
                for (const auto &a : c.tail) { cone.push_back(row_of(a, kk)); if (k < 0) k = kk; else if (kk >= 0 && kk != k) fail("a cone couples two nodes"); }
                if (k < 0) fail("a cone without variables");
                cones[k].push_back(cone);
            }
        }
        // the tables are per node: every node must carry the same rows
        t.lp = lp[0]; t.cones = cones[0];
        for (int k = 1; k < t.K; k++) {
            if (lp[k].size() != t.lp.size() || cones[k].size() != t.cones.size()) fail("the model's inequality rows differ from node to node");
            for (size_t r = 0; r < t.lp.size(); r++) if (!lp[k][r].same(t.lp[r])) fail("an LP row differs from node to node");
            for (size_t q = 0; q < t.cones.size(); q++) {
                if (cones[k][q].size() != t.cones[q].size()) fail("a cone differs from node to node");
                for (size_t r = 0; r < t.cones[q].size(); r++) if (!cones[k][q][r].same(t.cones[q][r])) fail("a cone row differs from node to node");
            }
        }
        // pins: present at every node / at the first / at the last node
        auto has = [](const std::vector<Pin> &l, const Pin &q) { for (auto &e : l) if (e.idx == q.idx && e.value.same(q.value)) return true; return false; };
        for (int k = 0; k < t.K; k++)
            for (const auto &q : pins[k]) {
                bool everywhere = true;
                for (int j = 0; j < t.K; j++) everywhere = everywhere && has(pins[j], q);
                std::vector<Pin> *dst = everywhere ? &t.pin_all : (k == 0 ? &t.pin_first : (k == t.K - 1 ? &t.pin_last : nullptr));
                if (!dst) fail("a variable is pinned at an interior node only");
                if (!has(*dst, q)) dst->push_back(q);
            }
        return t;
    }

private:
    cvx::OptimizationProblem::Var X_, U_;
    int K_ = 0;
    [[noreturn]] static void fail(const std::string &m) { throw std::runtime_error("scpp_plugin: cannot lower the recorded constraints: " + m); }
    bool in(const Region &r, const double *p) const { return r.base && p >= r.base && p < r.base + r.n; }
    // variable id -> (node, index in xi = [x ; u])
    bool decode(int id, int &k, int &idx) const
    {
        if (id >= X_.offset && id < X_.offset + X_.rows * X_.cols) { k = (id - X_.offset) / X_.rows; idx = (id - X_.offset) % X_.rows; return true; }
        if (id >= U_.offset && id < U_.offset + U_.rows * U_.cols) { k = (id - U_.offset) / U_.rows; idx = X_.rows + (id - U_.offset) % U_.rows; return true; }
        return false;
    }
    Source source_of(const cvx::Param &p, int k) const
    {
        Source s;
        if (!p.is_dynamic()) { s.kind = Source::LITERAL; s.lit = p.lit; return s; }
        s.scale = p.scale;
        if (in(constants, p.ptr)) { s.kind = Source::CONSTANT; s.index = int(p.ptr - constants.base); }
        else if (in(x_init, p.ptr)) { s.kind = Source::X_INIT; s.index = int(p.ptr - x_init.base); }
        else if (in(x_final, p.ptr)) { s.kind = Source::X_FINAL; s.index = int(p.ptr - x_final.base); }
        else if (in(node_array, p.ptr)) {
            const int off = int(p.ptr - node_array.base);
            if (k >= 0 && off / node_rows != k) fail("a per-node parameter is used at another node");
            s.kind = Source::NODE_ARRAY; s.index = off % node_rows;
        } else fail("a dynpar points outside the regions the model registered (constants, x_init, x_final, per-node array)");
        return s;
    }
    int k_of(const cvx::Affine &a) const
    {
        int k = -1;
        for (const auto &t : a.terms) { int kk, idx; if (!decode(t.var, kk, idx)) fail("a constraint uses a variable other than X / U"); if (k < 0) k = kk; else if (kk != k) fail("a row couples two nodes"); }
        if (k < 0) fail("a constraint without variables");
        return k;
    }
    // a (>= 0 or in a cone) as  h - sum c xi : terms with the same variable are merged when both coefficients are literals
    Row row_of(const cvx::Affine &a, int &k) const
    {
        Row r;
        k = -1;
        for (const auto &t : a.terms) { int kk, idx; if (!decode(t.var, kk, idx)) fail("a constraint uses a variable other than X / U"); if (k < 0) k = kk; else if (kk != k) fail("a row couples two nodes"); }
        for (const auto &t : a.terms) {
            int kk, idx; decode(t.var, kk, idx);
            Source s = source_of(-t.coef, k);
            bool merged = false;
            for (auto &e : r.c) if (e.first == idx && e.second.kind == Source::LITERAL && s.kind == Source::LITERAL) { e.second.lit += s.lit; merged = true; }
            if (!merged) r.c.push_back({idx, s});
        }
        r.c.erase(std::remove_if(r.c.begin(), r.c.end(), [](const std::pair<int, Source> &e) { return e.second.kind == Source::LITERAL && e.second.lit == 0.; }), r.c.end());
        Source h; h.kind = Source::LITERAL; h.lit = 0.;
        bool dyn = false;
        for (const auto &p : a.consts) {
            Source s = source_of(p, k);
            if (s.kind == Source::LITERAL) { if (dyn) { if (s.lit != 0.) fail("right-hand side mixes a parameter and a literal"); } else h.lit += s.lit; }
            else { if (dyn || h.lit != 0.) fail("right-hand side with more than one parameter"); h = s; dyn = true; }
        }
        r.h = h;
        return r;
    }
};

// ---- code generation: the table as the macros scpp_b200/csrc/models.cuh expects from a generated model description ------------
// constant slots: 0, 1, -1 are slots 0..2 (as in the hand-written tables); every other distinct coefficient source gets the next slot;
// RECIPE lists how the device fills them per instance:  {kind, index, scale or literal}
struct Emitted { std::string text; int nlp, ncone, ncr, maxdim, ncst; };
inline Emitted emit_inc(const StageTable &t, const std::string &NAME, int max_cst = 12)
{
    std::vector<Source> slots(3);
    slots[0].lit = 0.; slots[1].lit = 1.; slots[2].lit = -1.;
    auto slot_of = [&](const Source &s) {
        for (size_t i = 0; i < slots.size(); i++) if (slots[i].same(s)) return (int)i;
        slots.push_back(s);
        return (int)slots.size() - 1;
    };
    auto rowdesc = [&](const Row &r) {
        if (r.c.size() > 3) throw std::runtime_error("scpp_plugin: a row with more than 3 entries does not fit RowDesc");
        std::ostringstream o;
        int idx[3] = {0, 0, 0}, cs[3] = {0, 0, 0};
        for (size_t q = 0; q < r.c.size(); q++) {
            idx[q] = r.c[q].first;
            const Source &s = r.c[q].second;
            if (s.kind == Source::NODE_ARRAY) {
                // RowDesc: cs < 0 takes -tdir[-cs-1]; the recorded coefficient is c = scale * array, so scale must be -1 (n' T >= T_min)
                if (s.scale != -1.) throw std::runtime_error("scpp_plugin: per-node coefficient with an unsupported sign / scale");
                cs[q] = -(s.index + 1);
            } else cs[q] = slot_of(s);
        }
        o << "{" << r.c.size() << ", {" << idx[0] << ", " << idx[1] << ", " << idx[2] << "}, {" << cs[0] << ", " << cs[1] << ", " << cs[2] << "}, " << slot_of(r.h) << "}";
        return o.str();
    };
    std::ostringstream o;
    o << "// GENERATED by tools/gen_plugin.cpp from the model's addApplicationConstraints (recorded through include/scpp_cvx.hpp, lowered by\n"
         "// include/scpp_plugin.hpp).  Do not edit: re-run scpp_b200/build.py.\n";
    o << "#define " << NAME << "_NLP " << t.lp.size() << "\n#define " << NAME << "_NCONE " << t.cones.size() << "\n#define " << NAME << "_NCR " << t.cone_rows()
      << "\n#define " << NAME << "_MAXDIM " << t.max_cone_dim() << "\n";
    o << "#define " << NAME << "_CONE_DIMS {";
    for (size_t q = 0; q < t.cones.size(); q++) o << (q ? ", " : "") << t.cones[q].size();
    o << "}\n#define " << NAME << "_CONE_OFFS {";
    { int off = 0; for (size_t q = 0; q < t.cones.size(); q++) { o << (q ? ", " : "") << off; off += (int)t.cones[q].size(); } }
    o << "}\n#define " << NAME << "_ROWS { \\\n";
    for (const auto &r : t.lp) o << "        " << rowdesc(r) << ", \\\n";
    for (const auto &c : t.cones) for (const auto &r : c) o << "        " << rowdesc(r) << ", \\\n";
    o << "    }\n";
    if ((int)slots.size() > max_cst) throw std::runtime_error("scpp_plugin: more coefficient slots than MAX_CST");
    // recipe: kind 0 literal, 1 constants[index] * scale, 2 x_init[index] * scale, 3 x_final[index] * scale (the scaled boundary states)
    o << "#define " << NAME << "_NCST " << slots.size() << "\n#define " << NAME << "_CST_RECIPE { \\\n";
    for (const auto &s : slots) {
        if (s.kind == Source::LITERAL) o << "        {0, 0, " << s.lit << "}, \\\n";
        else if (s.kind == Source::CONSTANT) o << "        {1, " << s.index << ", " << s.scale << "}, \\\n";
        else if (s.kind == Source::X_INIT) o << "        {2, " << s.index << ", " << s.scale << "}, \\\n";
        else if (s.kind == Source::X_FINAL) o << "        {3, " << s.index << ", " << s.scale << "}, \\\n";      // e.g. m_dry = x_final(0), rocketQuat.cpp:93
        else throw std::runtime_error("scpp_plugin: unsupported coefficient source");
    }
    o << "    }\n";
    auto pins = [&](const char *what, const std::vector<Pin> &l) {
        o << "#define " << NAME << "_PIN_" << what << " { \\\n";
        for (const auto &p : l) {
            const int kind = p.value.kind == Source::LITERAL ? 0 : (p.value.kind == Source::X_INIT ? 1 : (p.value.kind == Source::X_FINAL ? 2 : -1));
            if (kind < 0) throw std::runtime_error("scpp_plugin: a variable pinned to a model constant is not supported");
            o << "        {" << p.idx << ", " << kind << ", " << p.value.index << ", " << (kind == 0 ? p.value.lit : p.value.scale) << "}, \\\n";
        }
        o << "        {-1, 0, 0, 0.} \\\n    }\n";
    };
    pins("ALL", t.pin_all); pins("FIRST", t.pin_first); pins("LAST", t.pin_last);
    Emitted e; e.text = o.str(); e.nlp = (int)t.lp.size(); e.ncone = (int)t.cones.size(); e.ncr = t.cone_rows(); e.maxdim = t.max_cone_dim(); e.ncst = (int)slots.size();
    return e;
}

} // namespace scpp_plugin
