// Dev tool (CPU only; links the ORACLE, so it is test infrastructure and never part of the product).
// SURVEY.md section 7 step 5 / VERDICT r01 item 3: for every alignment of oracle-driven progressive
// builds, compare the path of the dense DP model (oracle/dp_model.hpp = what the DP kernels compute)
// with the path of the A* restatement and classify the first divergence; then evaluate on-device
// detectors that would let an alignment skip the order-exact search.
//
//   g++ -O2 -std=c++17 -o /tmp/tie_study tools/tie_study.cpp -x c poasta_b200/csrc/synth.c && /tmp/tie_study 60 1000 0.05 20
#include <cstdio>
#include <cstdlib>
#include <map>
#include <unordered_map>
#include <unordered_set>

#include "../oracle/astar_oracle.hpp"
#include "../oracle/dp_model.hpp"
#include "../oracle/poa_oracle.hpp"

extern "C" size_t synth_family(uint64_t seed, uint32_t n_seqs, uint32_t root_len, double rate, const char *prefix, char *out,
                               size_t cap);

using namespace oracle;

static inline uint64_t key(uint32_t n, uint32_t o, uint32_t s) { return ((uint64_t)n << 34) | ((uint64_t)o << 2) | s; }

struct Step {
    uint32_t n, o;
    int s;
};

// The reference's backtrace (gap_affine.rs:804-915) over an arbitrary score lookup; records every cell it stands on and, for
// each step, how many candidates qualified (in priority order) -- `ncand` > 1 means the table held a tie there.
template <class Val> struct Walker {
    const poasta_graph_view &g;
    const uint8_t *q;
    uint32_t x, o, e;
    Val val;  // (node, off, state) -> score or 0xFFFFFFFF
    // enumerates qualifying candidates of a cell in priority order
    void candidates(uint32_t node, uint32_t off, int st, std::vector<Step> &out) const {
        out.clear();
        const uint32_t cur = val(node, off, st);
        if (cur == 0xFFFFFFFFu) return;
        auto eq = [](uint32_t ps, uint32_t t) { return ps != 0xFFFFFFFFu && ps == t; };
        const uint32_t *pb = g.pred + g.pred_off[node], *pe = g.pred + g.pred_off[node + 1];
        if (st == 0) {
            if (off > 0) {
                const bool me = node == g.end_node || g.symbol[node] == q[off - 1];
                const uint32_t po = node == g.end_node ? off : off - 1;
                for (const uint32_t *p = pb; p != pe; ++p) {
                    uint32_t ps = val(*p, po, 0);
                    if ((me && eq(ps, cur)) || (!me && eq(ps, cur - x))) out.push_back({*p, po, 0});
                }
            }
            if (eq(val(node, off, 1), cur)) out.push_back({node, off, 1});
            if (eq(val(node, off, 2), cur)) out.push_back({node, off, 2});
        } else if (st == 1) {
            for (const uint32_t *p = pb; p != pe; ++p)
                if (eq(val(*p, off, 0), cur - o - e)) out.push_back({*p, off, 0});
            for (const uint32_t *p = pb; p != pe; ++p)
                if (eq(val(*p, off, 1), cur - e)) out.push_back({*p, off, 1});
        } else if (off > 0) {
            if (eq(val(node, off - 1, 0), cur - o - e)) out.push_back({node, off - 1, 0});
            if (eq(val(node, off - 1, 2), cur - e)) out.push_back({node, off - 1, 0});  // sic
        }
    }
};

struct Counters {
    uint64_t n = 0, same_score = 0, same_path = 0;
    uint64_t unique = 0;            // DP walk never saw two qualifying candidates
    uint64_t unique_and_same = 0;
    std::map<std::string, uint64_t> cls;
    uint64_t tie_steps = 0, steps = 0;
    // detector experiments: an alignment "passes" when every tie on the DP walk is certified
    uint64_t pass_det[8] = {0}, pass_det_same[8] = {0};
};

int main(int argc, char **argv) {
    uint32_t nfam = argc > 1 ? atoi(argv[1]) : 20;
    uint32_t L = argc > 2 ? atoi(argv[2]) : 300;
    double rate = argc > 3 ? atof(argv[3]) : 0.05;
    uint32_t nseq = argc > 4 ? atoi(argv[4]) : 10;
    int verbose = argc > 5 ? atoi(argv[5]) : 0;
    const uint32_t X = 4, O = 6, E = 2;
    std::vector<char> buf(64 << 20);
    Counters c;
    for (uint32_t f = 0; f < nfam; ++f) {
        size_t len = synth_family(0xC200000000ull + f, nseq, L, rate, "f", buf.data(), buf.size());
        auto recs = poa::read_fasta(buf.data(), len);
        poa::POAGraph g;
        for (size_t i = 0; i < recs.size(); ++i) {
            const uint8_t *s = (const uint8_t *)recs[i].seq.data();
            const uint32_t Q = (uint32_t)recs[i].seq.size();
            const std::vector<uint64_t> ones(Q, 1);
            if (g.is_empty()) {
                g.add_alignment_with_weights(recs[i].name, recs[i].seq, nullptr, ones);
                continue;
            }
            const poa::Flat flat(g);
            poasta_graph_view view = flat.view;
            AlignOptions opt;
            std::vector<uint32_t> dump, pruned, pops;
            opt.dump = &dump;
            opt.pruned_log = &pruned;
            opt.pop_log = &pops;
            AstarResult r = astar_align(view, s, Q, opt);
            if (r.status) {
                fprintf(stderr, "oracle panic %d\n", r.status);
                return 1;
            }
            std::unordered_map<uint64_t, uint32_t> A;
            A.reserve(dump.size() / 2);
            for (size_t k = 0; k + 3 < dump.size(); k += 4) A[key(dump[k], dump[k + 1], dump[k + 2])] = dump[k + 3];
            std::unordered_set<uint64_t> P, X_;
            for (size_t k = 0; k + 2 < pruned.size(); k += 3) P.insert(key(pruned[k], pruned[k + 1], pruned[k + 2]));
            for (size_t k = 0; k + 3 < pops.size(); k += 4) X_.insert(key(pops[k], pops[k + 1], pops[k + 2]));

            dpmodel::Tables T;
            std::vector<uint32_t> drp, dqp;
            uint32_t dsc = dpmodel::align_global(view, s, Q, X, O, E, drp, dqp, &T);
            ++c.n;
            c.same_score += dsc == r.score;
            const bool same = drp == r.rpos && dqp == r.qpos;
            c.same_path += same;

            auto dpval = [&](uint32_t n, uint32_t o, int st) -> uint32_t {
                uint32_t v = st == 0 ? T.m(n, o) : (st == 1 ? T.d(n, o) : T.i(n, o));
                return v >= dpmodel::INF ? 0xFFFFFFFFu : v;
            };
            auto aval = [&](uint32_t n, uint32_t o, int st) -> uint32_t {
                auto it = A.find(key(n, o, st));
                return it == A.end() ? 0xFFFFFFFFu : it->second;
            };
            Walker<decltype(dpval)> wd{view, s, X, O, E, dpval};
            Walker<decltype(aval)> wa{view, s, X, O, E, aval};

            Graph og(view);
            BubbleIndex bi(og);
            Heuristic heur;
            heur.kind = 1, heur.costs = Costs{4, 6, 2}, heur.bubbles = &bi, heur.seq_len = Q;

            // walk the DP path; at every step compare the candidate sets of both tables
            std::vector<Step> cd, ca;
            Step cur{view.end_node, Q, 0};
            bool unique = true, diverged = false;
            bool det_ok[8] = {true, true, true, true, true, true, true, true};
            std::string cls;
            while (true) {
                wd.candidates(cur.n, cur.o, cur.s, cd);
                if (cd.empty()) break;
                ++c.steps;
                if (cd.size() > 1) {
                    unique = false;
                    ++c.tie_steps;
                    // detectors, evaluated on the first-priority candidate c1 of a tie
                    const Step &c1 = cd[0];
                    const uint32_t g1 = dpval(c1.n, c1.o, c1.s);
                    const uint32_t f1 = g1 + (uint32_t)heur.h(c1.n, c1.o, (State)c1.s);
                    // det0: f(c1) < C*
                    if (!(f1 < dsc)) det_ok[0] = false;
                }
                if (!diverged) {
                    wa.candidates(cur.n, cur.o, cur.s, ca);
                    if (ca.empty() || ca[0].n != cd[0].n || ca[0].o != cd[0].o || ca[0].s != cd[0].s) {
                        diverged = true;
                        const Step &c1 = cd[0];
                        const uint32_t g1 = dpval(c1.n, c1.o, c1.s);
                        const uint32_t a1 = aval(c1.n, c1.o, c1.s);
                        const uint32_t f1 = g1 + (uint32_t)heur.h(c1.n, c1.o, (State)c1.s);
                        static const char *SN[3] = {"M", "D", "I"};
                        cls = std::string("cur") + SN[cur.s] + "->dp" + SN[c1.s] + (ca.empty() ? "/a-" : std::string("/a") + SN[ca[0].s]);
                        cls += a1 == 0xFFFFFFFFu ? " unrec" : (a1 > g1 ? " worse" : " ??");
                        cls += f1 < dsc ? " f<C" : (f1 == dsc ? " f=C" : " f>C");
                        // why: walk the DP path back from c1 until a cell A* recorded at g*; was that cell expanded / pruned?
                        Step w = c1;
                        std::vector<Step> tmp;
                        int hops = 0;
                        while (true) {
                            wd.candidates(w.n, w.o, w.s, tmp);
                            if (tmp.empty()) break;
                            // any optimal predecessor that A* recorded at g*?
                            bool found = false;
                            for (auto &p : tmp) {
                                if (aval(p.n, p.o, p.s) == dpval(p.n, p.o, p.s)) {
                                    found = true;
                                    const uint64_t k = key(p.n, p.o, p.s);
                                    cls += std::string(" via") + SN[p.s];
                                    cls += P.count(k) ? " PRUNED" : (X_.count(k) ? " expanded" : " notpopped");
                                    const uint32_t fp = dpval(p.n, p.o, p.s) + (uint32_t)heur.h(p.n, p.o, (State)p.s);
                                    cls += fp < dsc ? "" : " (f>=C)";
                                    break;
                                }
                            }
                            if (found) break;
                            w = tmp[0];
                            ++hops;
                        }
                        cls += " hops" + std::to_string(hops > 3 ? 4 : hops);
                        if (verbose && c.cls[cls] < (uint64_t)verbose)
                        {
                            fprintf(stderr, "fam %u seq %zu V=%u Q=%u C*=%u at (%u,%u,%s) score %u: %s\n", f, i, view.n_nodes, Q, dsc, cur.n,
                                    cur.o, SN[cur.s], dpval(cur.n, cur.o, cur.s), cls.c_str());
                            auto show = [&](const char *tag, const std::vector<uint32_t> &rp, const std::vector<uint32_t> &qp) {
                                // pairs whose query position lies in [cur.o-12, cur.o+3]
                                std::string l1, l2;
                                for (size_t k = 0; k < rp.size(); ++k) {
                                    uint32_t qq = qp[k];
                                    if (qq == NONE) {  // deletion: place by neighbours
                                        size_t k2 = k;
                                        while (k2 < qp.size() && qp[k2] == NONE) ++k2;
                                        qq = k2 < qp.size() ? qp[k2] : Q;
                                    }
                                    if (qq + 14 < cur.o || qq > cur.o + 4) continue;
                                    l1 += rp[k] == NONE ? '-' : (char)view.symbol[rp[k]];
                                    l2 += qp[k] == NONE ? '-' : (char)s[qp[k]];
                                }
                                fprintf(stderr, "   %s ref %s\n   %s qry %s\n", tag, l1.c_str(), tag, l2.c_str());
                            };
                            show("A*", r.rpos, r.qpos);
                            show("DP", drp, dqp);
                        }
                    }
                }
                if (cd[0].n == view.start_node && !(cur.s == 0 && cd[0].s != 0)) break;
                cur = cd[0];
            }
            if (same != !diverged && verbose) fprintf(stderr, "note: walk/path disagreement fam %u seq %zu same=%d diverged=%d\n", f, i, same, diverged);
            c.unique += unique;
            c.unique_and_same += unique && same;
            if (diverged) ++c.cls[cls];
            for (int d = 0; d < 8; ++d) {
                c.pass_det[d] += det_ok[d];
                c.pass_det_same[d] += det_ok[d] && same;
            }

            std::vector<poa::Pair> aln(r.rpos.size());
            for (size_t k = 0; k < aln.size(); ++k) aln[k] = {r.rpos[k], r.qpos[k]};
            g.add_alignment_with_weights(recs[i].name, recs[i].seq, &aln, ones);
        }
    }
    printf("alignments %lu  same_score %lu  same_path %lu (%.1f%%)\n", c.n, c.same_score, c.same_path, 100.0 * c.same_path / c.n);
    printf("tie-free DP walks %lu (%.1f%%), of which same path %lu; tie steps %.2f per alignment (%.2f%% of steps)\n", c.unique,
           100.0 * c.unique / c.n, c.unique_and_same, (double)c.tie_steps / c.n, 100.0 * c.tie_steps / c.steps);
    for (int d = 0; d < 1; ++d)
        printf("detector %d: pass %lu (%.1f%%), of which same path %lu -> unsound on %lu\n", d, c.pass_det[d], 100.0 * c.pass_det[d] / c.n,
               c.pass_det_same[d], c.pass_det[d] - c.pass_det_same[d]);
    printf("first-divergence classes:\n");
    std::vector<std::pair<uint64_t, std::string>> v;
    for (auto &kv : c.cls) v.push_back({kv.second, kv.first});
    std::sort(v.rbegin(), v.rend());
    for (auto &p : v) printf("  %6lu  %s\n", p.first, p.second.c_str());
    return 0;
}
