#include "plan.h"

#include <algorithm>
#include <cstring>
#include <functional>
#include <numeric>
#include <unordered_map>

#include <cstdlib>

#include "hand_indexer.h"
#include "poker.h"

namespace rs {

namespace {

struct Builder {
    const rs_tree* t;
    Plan* P;
    std::string err;
    uint32_t max_round = 0;

    bool fail(const std::string& m) {
        if (err.empty()) err = m;
        return false;
    }

    // DFS copy of the reference tree into PNodes. Returns PNode id or -1.
    int32_t copy_node(uint32_t id, uint32_t round_k, int depth) {
        if (id >= t->n_nodes) return fail("child id out of range"), -1;
        if (depth > 4096) return fail("tree too deep / cyclic"), -1;
        uint32_t c0 = t->child_offset[id], c1 = t->child_offset[id + 1];
        max_round = std::max(max_round, round_k);
        switch (t->type[id]) {
            case RS_NODE_ACTION: {
                if (c1 <= c0) return fail("action node without children"), -1;
                if (t->player[id] > 1) return fail("player must be 0 or 1"), -1;
                if (t->round_idx[id] != round_k) return fail("ActionNode.round_idx does not match chance depth"), -1;
                int32_t me = int32_t(P->nodes.size());
                P->nodes.emplace_back();
                {
                    PNode& n = P->nodes[me];
                    n.kind = PK_ACTION;
                    n.player = t->player[id];
                    n.round_k = uint8_t(round_k);
                    n.an_index = t->an_index[id];
                    n.src_node = int32_t(id);
                }
                std::vector<int32_t> ch;
                for (uint32_t i = c0; i < c1; ++i) {
                    int32_t c = copy_node(t->children[i], round_k, depth + 1);
                    if (c < 0) return -1;
                    ch.push_back(c);
                }
                P->nodes[me].children = ch;
                return me;
            }
            case RS_NODE_TERMINAL: {
                int32_t me = int32_t(P->nodes.size());
                P->nodes.emplace_back();
                PNode n;
                n.round_k = uint8_t(round_k);
                n.value = t->value[id];
                n.last_to_act = t->last_to_act[id];
                n.src_node = int32_t(id);
                uint32_t abs_round = P->first_round + round_k;
                if (t->ttype[id] == RS_TERM_UNCONTESTED) {
                    n.kind = PK_FOLD;
                    P->nodes[me] = n;
                } else if (abs_round >= RS_ROUND_RIVER) {
                    n.kind = PK_SHOWDOWN;
                    P->nodes[me] = n;
                } else {
                    // ALLIN (or showdown) before the river: expected showdown over run-outs.
                    // The reference's cfr() evaluates these with undealt cards (cfr.rs:544-556,
                    // a latent bug); mccfr() is right in expectation because the whole board is
                    // pre-sampled (cfr.rs:114-122).  We insert the run-out chance nodes.
                    n.kind = PK_CHANCE;
                    P->nodes[me] = n;
                    int32_t cur = me;
                    uint32_t rk = round_k;
                    while (P->first_round + rk + 1 < RS_ROUND_RIVER) {
                        int32_t nx = int32_t(P->nodes.size());
                        P->nodes.emplace_back();
                        PNode c;
                        c.kind = PK_CHANCE;
                        c.round_k = uint8_t(rk + 1);
                        c.value = n.value;
                        P->nodes[nx] = c;
                        P->nodes[cur].children = {nx};
                        cur = nx;
                        rk++;
                    }
                    int32_t sdn = int32_t(P->nodes.size());
                    P->nodes.emplace_back();
                    PNode s;
                    s.kind = PK_SHOWDOWN;
                    s.round_k = uint8_t(rk + 1);
                    s.value = n.value;
                    s.last_to_act = n.last_to_act;
                    P->nodes[sdn] = s;
                    P->nodes[cur].children = {sdn};
                    max_round = std::max(max_round, rk + 1);
                }
                return me;
            }
            case RS_NODE_PUBLIC_CHANCE: {
                if (c1 != c0 + 1) return fail("public chance node needs exactly one child"), -1;
                if (P->first_round + round_k >= RS_ROUND_RIVER) return fail("chance node after the river"), -1;
                int32_t me = int32_t(P->nodes.size());
                P->nodes.emplace_back();
                P->nodes[me].kind = PK_CHANCE;
                P->nodes[me].round_k = uint8_t(round_k);
                P->nodes[me].src_node = int32_t(id);
                int32_t c = copy_node(t->children[c0], round_k + 1, depth + 1);
                if (c < 0) return -1;
                P->nodes[me].children = {c};
                return me;
            }
            default:
                return fail("unexpected node type below the root"), -1;
        }
    }

    // carve the PNode tree into per-round segments
    void make_segment(uint32_t k, int32_t root) {
        uint32_t sid = uint32_t(P->segs[k].size());
        P->segs[k].emplace_back();
        P->segs[k][sid].root = root;
        std::vector<int32_t> leaves;
        std::function<void(int32_t, int)> walk = [&](int32_t id, int depth) {
            PNode& n = P->nodes[id];
            n.depth = depth;
            if (n.kind == PK_CHANCE) {
                leaves.push_back(id);
                return;
            }
            for (int32_t c : n.children) {
                P->nodes[c].parent = id;
                walk(c, depth + 1);
            }
        };
        walk(root, 0);
        P->segs[k][sid].leaves = leaves;
        for (int32_t leaf : leaves) {
            P->nodes[leaf].leaf_id = int32_t(P->segs[k + 1].size());
            make_segment(k + 1, P->nodes[leaf].children[0]);
        }
    }
};

// ---- task-graph generation (tasks.h) -----------------------------------------------------------

struct TaskGen {
    Plan* P;
    int trav;
    TaskList tl;
    StreetPlan sp;
    std::string err;

    struct RSrc {
        int32_t buf = RIN_INITIAL;
        uint8_t parent_round = 0;
    };
    std::vector<RSrc> rsrc;                 // per PNode: where its incoming opponent reach lives
    std::vector<int32_t> cbuf, tbuf;        // per PNode: value buffer / terminal-partial buffer
    std::vector<int32_t> down_task, up_task;  // per PNode: node-task index
    std::vector<int32_t> terms_task, terms_buf;  // per PNode (traverser nodes of chain rounds): TK_TRAV_TERMS task, its first value buffer
    std::vector<int32_t> rbuf_producer[3];  // reach buffer id -> node-task index writing it
    std::vector<int32_t> leaf_rbuf[3];      // leaf id -> reach buffer (of the leaf's round) the child street reads
    std::vector<int32_t> leaf_gather[3];    // leaf id -> node-task index of its gather
    std::vector<int32_t> segroot_cbuf[3], segroot_task[3];

    struct Pending {
        NodeTask t;
        int32_t pnode;
        int depth;
    };

    // a round with one or two local boards is a chain of dependent tasks (tasks.h: TK_TRAV_TERMS)
    // RS_CHAIN_BOARDS overrides the board count up to which a round is treated as a chain (tuning)
    bool chain_round(uint32_t k) const {
        static const uint32_t limit = [] {
            const char* e = getenv("RS_CHAIN_BOARDS");
            return e ? uint32_t(atoi(e)) : 2u;
        }();
        return P->boards_local(k) <= limit && !(P->flags & RS_FLAG_NO_CHAIN_SPLIT);
    }

    int32_t new_rbuf(uint32_t k) {
        rbuf_producer[k].push_back(-1);
        return int32_t(tl.n_rbuf[k]++);
    }
    int32_t new_cbuf(uint32_t k) { return int32_t(tl.n_cbuf[k]++); }

    static NodeTask blank(uint8_t kind, uint32_t k) {
        NodeTask t;
        std::memset(&t, 0, sizeof(t));
        t.kind = kind;
        t.round_k = uint8_t(k);
        t.r_in = RIN_INITIAL;
        t.out = -1;
        t.aux = -1;
        for (int i = 0; i < MAX_TASK_DEPS; ++i) t.dep[i] = -1;
        return t;
    }
    bool add_dep(NodeTask& t, int32_t task, uint8_t kind) {
        if (task < 0) return true;
        if (t.n_dep >= MAX_TASK_DEPS) {
            err = "too many dependencies for one task";
            return false;
        }
        t.dep[t.n_dep] = task;
        t.dep_kind[t.n_dep] = kind;
        t.n_dep++;
        return true;
    }
    bool add_rin_dep(NodeTask& t, const RSrc& r, uint32_t k) {
        t.r_in = r.buf;
        t.rin_parent_round = r.parent_round;
        if (r.buf == RIN_INITIAL) return true;
        if (r.parent_round) return add_dep(t, rbuf_producer[k - 1][r.buf], DK_PARENT_BOARD);
        return add_dep(t, rbuf_producer[k][r.buf], DK_SAME_BOARD);
    }
    uint32_t emit(NodeTask t) {
        t.first = tl.n_tickets;
        t.count = P->boards_local(t.round_k);
        tl.n_tickets += t.count;
        tl.tasks.push_back(t);
        return uint32_t(tl.tasks.size() - 1);
    }

    // assign reach sources + value buffers for one street segment (DFS), collect its nodes
    void assign(uint32_t k, int32_t id, const RSrc& r, std::vector<int32_t>& order) {
        PNode& n = P->nodes[id];
        rsrc[id] = r;
        order.push_back(id);
        if (n.kind != PK_ACTION) return;
        cbuf[id] = new_cbuf(k);
        const bool opp = (n.player != trav);
        bool has_term = false;
        for (int32_t c : n.children) {
            const PNode& cn = P->nodes[c];
            if (cn.kind == PK_FOLD || cn.kind == PK_SHOWDOWN) {
                has_term = true;
                rsrc[c] = r;  // unused for opp parents (staged in shared memory)
                continue;
            }
            RSrc cr = r;
            if (opp) {
                cr.buf = new_rbuf(k);
                cr.parent_round = 0;
            }
            assign(k, c, cr, order);
        }
        if (opp && has_term) tbuf[id] = new_cbuf(k);
    }

    bool run() {
        const size_t N = P->nodes.size();
        rsrc.assign(N, RSrc());
        cbuf.assign(N, -1);
        tbuf.assign(N, -1);
        down_task.assign(N, -1);
        up_task.assign(N, -1);
        terms_task.assign(N, -1);
        terms_buf.assign(N, -1);
        const uint32_t R = P->n_rounds;
        std::vector<std::vector<int32_t>> seg_nodes[3];
        for (uint32_t k = 0; k < R; ++k) {
            leaf_rbuf[k].assign(k + 1 < R ? P->segs[k + 1].size() : 0, -1);
            leaf_gather[k].assign(leaf_rbuf[k].size(), -1);
            segroot_cbuf[k].assign(P->segs[k].size(), -1);
            segroot_task[k].assign(P->segs[k].size(), -1);
        }
        // ---------------- down tasks, street by street ----------------
        for (uint32_t k = 0; k < R; ++k) {
            std::vector<Pending> pend;
            seg_nodes[k].resize(P->segs[k].size());
            for (uint32_t s = 0; s < P->segs[k].size(); ++s) {
                RSrc root;
                if (k > 0) {
                    root.buf = leaf_rbuf[k - 1][s];
                    root.parent_round = 1;
                    if (root.buf < 0) {
                        err = "internal: chance leaf without a reach buffer";
                        return false;
                    }
                }
                const int32_t rootid = P->segs[k][s].root;
                assign(k, rootid, root, seg_nodes[k][s]);
                if (k > 0) cbuf[rootid] = int32_t(tl.n_sbuf[k]++);  // street roots below the first round: scatter pool
                else if (P->nodes[rootid].kind != PK_ACTION) cbuf[rootid] = new_cbuf(k);
                segroot_cbuf[k][s] = cbuf[rootid];
            }
            // chance leaves whose inherited source is not a buffer of this round get a copy task
            for (uint32_t s = 0; s < P->segs[k].size(); ++s)
                for (int32_t id : seg_nodes[k][s]) {
                    PNode& n = P->nodes[id];
                    if (n.kind == PK_CHANCE) {
                        RSrc r = rsrc[id];
                        if (r.parent_round || r.buf == RIN_INITIAL) {
                            NodeTask t = blank(TK_CHANCE_DOWN, k);
                            RSrc keep = r;
                            const int32_t nb = new_rbuf(k);
                            t.aux = nb;
                            // dependency is added at emission (producer index known by then for parent rounds)
                            pend.push_back({t, id, n.depth});
                            pend.back().t.seg = s;
                            rsrc[id].buf = nb;
                            rsrc[id].parent_round = 0;
                            // stash the original source in the task
                            pend.back().t.r_in = keep.buf;
                            pend.back().t.rin_parent_round = keep.parent_round;
                        }
                        leaf_rbuf[k][n.leaf_id] = rsrc[id].buf;
                    } else if (n.kind == PK_ACTION && n.player != trav) {
                        NodeTask t = blank(TK_DOWN, k);
                        t.n_act = uint8_t(n.children.size());
                        t.cum_a = n.cum_a;
                        t.an_index = n.an_index;
                        t.out = tbuf[id];
                        uint32_t nterm = 0;
                        for (size_t a = 0; a < n.children.size(); ++a) {
                            const int32_t c = n.children[a];
                            const PNode& cn = P->nodes[c];
                            TaskChild& tc = t.child[a];
                            tc.buf = -1;
                            if (cn.kind == PK_FOLD) {
                                tc.kind = CK_FOLD;
                                tc.coef = (trav == cn.last_to_act) ? -float(cn.value) : float(cn.value);  // cfr.rs:525-531
                                nterm++;
                            } else if (cn.kind == PK_SHOWDOWN) {
                                tc.kind = CK_SHOWDOWN;
                                tc.coef = float(cn.value);  // cfr.rs:532-543
                                nterm++;
                            } else {
                                tc.kind = cn.kind == PK_CHANCE ? CK_CHANCE : CK_ACTION;
                                tc.buf = rsrc[c].buf;
                            }
                        }
                        tl.max_terminal = std::max(tl.max_terminal, nterm);
                        if (chain_round(k) && nterm > 0 && nterm < n.children.size()) {
                            // chain round: the next level only waits for the child reach, so that goes first and
                            // alone; a second task values the terminal children
                            NodeTask reach = t;
                            reach.out = -1;
                            pend.push_back({reach, id, n.depth});
                            pend.back().t.seg = s;
                            for (size_t a = 0; a < n.children.size(); ++a)
                                if (t.child[a].kind == CK_ACTION || t.child[a].kind == CK_CHANCE) t.child[a].buf = -1;  // not written again
                        }
                        pend.push_back({t, id, n.depth});
                            pend.back().t.seg = s;
                    } else if (n.kind == PK_ACTION && n.player == trav && chain_round(k)) {
                        // chain round: scan + per-hand terms of the traverser node as soon as its reach exists
                        NodeTask t = blank(TK_TRAV_TERMS, k);
                        t.n_act = uint8_t(n.children.size());
                        t.an_index = n.an_index;
                        t.out = new_cbuf(k);  // mass; the showdown terms go to the next buffer
                        (void)new_cbuf(k);
                        for (size_t a = 0; a < n.children.size(); ++a) {
                            const PNode& cn = P->nodes[n.children[a]];
                            t.child[a].buf = -1;
                            t.child[a].kind = cn.kind == PK_FOLD ? CK_FOLD : (cn.kind == PK_SHOWDOWN ? CK_SHOWDOWN : CK_VALUE);
                        }
                        terms_buf[id] = t.out;
                        pend.push_back({t, id, n.depth});
                            pend.back().t.seg = s;
                    }
                }
            std::stable_sort(pend.begin(), pend.end(), [](const Pending& a, const Pending& b) { return a.depth < b.depth; });
            for (Pending& pe : pend) {
                NodeTask t = pe.t;
                if (t.kind == TK_CHANCE_DOWN) {
                    RSrc src;
                    src.buf = t.r_in;
                    src.parent_round = t.rin_parent_round;
                    if (!add_rin_dep(t, src, k)) return false;
                    const uint32_t ti = emit(t);
                    rbuf_producer[k][t.aux] = int32_t(ti);
                    down_task[pe.pnode] = int32_t(ti);
                } else if (t.kind == TK_TRAV_TERMS) {
                    if (!add_rin_dep(t, rsrc[pe.pnode], k)) return false;
                    terms_task[pe.pnode] = int32_t(emit(t));
                } else {
                    if (!add_rin_dep(t, rsrc[pe.pnode], k)) return false;
                    const uint32_t ti = emit(t);
                    down_task[pe.pnode] = int32_t(ti);  // of a split node: the task emitted last, the one with the terminal values
                    for (int a = 0; a < t.n_act; ++a)
                        if (t.child[a].buf >= 0) rbuf_producer[k][t.child[a].buf] = int32_t(ti);
                }
            }
        }
        // ---------------- up tasks, deepest street first ----------------
        // Opponent nodes below a traverser node get no task of their own: the parent sums their terminal
        // partial and their children's values itself (one hop less per tree level).  Opponent nodes that
        // are street roots keep a TK_UP_OPP task because the gather / root read-out needs their value.
        tl.phase_cut = 0;
        for (int k = int(R) - 1; k >= 0; --k) {
            std::vector<Pending> pend;
            for (uint32_t s = 0; s < P->segs[k].size(); ++s)
                for (int32_t id : seg_nodes[k][s]) {
                    PNode& n = P->nodes[id];
                    const bool is_root = (id == P->segs[k][s].root);
                    if (n.kind == PK_SHOWDOWN && is_root) {
                        NodeTask t = blank(TK_ROOT_SHOWDOWN, uint32_t(k));
                        t.out = cbuf[id];
                        t.child[0].kind = CK_SHOWDOWN;
                        t.child[0].coef = float(n.value);
                        pend.push_back({t, id, n.depth});
                            pend.back().t.seg = s;
                    } else if (n.kind == PK_CHANCE && is_root) {
                        NodeTask t = blank(TK_CHANCE_UP, uint32_t(k));
                        t.out = cbuf[id];
                        t.aux = n.leaf_id;
                        pend.push_back({t, id, n.depth});
                            pend.back().t.seg = s;
                    } else if (n.kind == PK_ACTION) {
                        const bool opp = (n.player != trav);
                        if (opp && !is_root && P->nodes[n.parent].kind == PK_ACTION && P->nodes[n.parent].player == trav) continue;
                        NodeTask t = blank(opp ? TK_UP_OPP : TK_UP_TRAV, uint32_t(k));
                        t.n_act = uint8_t(n.children.size());
                        t.cum_a = n.cum_a;
                        t.an_index = n.an_index;
                        t.out = cbuf[id];
                        if (!opp) tl.max_children = std::max<uint32_t>(tl.max_children, t.n_act);
                        pend.push_back({t, id, n.depth});
                            pend.back().t.seg = s;
                    }
                }
            std::stable_sort(pend.begin(), pend.end(), [](const Pending& a, const Pending& b) { return a.depth > b.depth; });
            // value sources of node `c` seen from its parent: c's own value buffer if c has an up task, else
            // (opponent node folded into the parent) its terminal partial + its children's sources
            std::function<bool(int32_t, std::vector<TaskSrc>&)> collect = [&](int32_t c, std::vector<TaskSrc>& out) -> bool {
                const PNode& cn = P->nodes[c];
                if (cn.kind == PK_CHANCE) {
                    out.push_back(TaskSrc{cn.leaf_id, leaf_gather[k][cn.leaf_id], SK_GATHERED, {0, 0, 0}});
                    return true;
                }
                if (cn.kind != PK_ACTION) return true;  // terminals are inside the terminal partial
                if (up_task[c] >= 0) {
                    out.push_back(TaskSrc{cbuf[c], up_task[c], SK_CBUF, {0, 0, 0}});
                    return true;
                }
                // folded opponent node
                if (tbuf[c] >= 0) out.push_back(TaskSrc{tbuf[c], down_task[c], SK_CBUF, {0, 0, 0}});
                for (int32_t gc : cn.children)
                    if (!collect(gc, out)) return false;
                return true;
            };
            for (Pending& pe : pend) {
                NodeTask t = pe.t;
                const PNode& n = P->nodes[pe.pnode];
                if (t.kind == TK_ROOT_SHOWDOWN) {
                    if (!add_rin_dep(t, rsrc[pe.pnode], uint32_t(k))) return false;
                } else if (t.kind == TK_CHANCE_UP) {
                    if (!add_dep(t, leaf_gather[k][n.leaf_id], DK_SAME_BOARD)) return false;
                } else {
                    if (t.kind == TK_UP_TRAV) {
                        if (terms_task[pe.pnode] >= 0) {  // chain round: the scan and the terms were done by TK_TRAV_TERMS
                            t.pre_terms = 1;
                            t.aux = terms_buf[pe.pnode];
                            if (!add_dep(t, terms_task[pe.pnode], DK_SAME_BOARD)) return false;
                        } else if (!add_rin_dep(t, rsrc[pe.pnode], uint32_t(k))) {
                            return false;
                        }
                    } else {
                        t.r_in = RIN_INITIAL;
                        t.aux = tbuf[pe.pnode];
                        if (!add_dep(t, down_task[pe.pnode], DK_SAME_BOARD)) return false;
                    }
                    t.src_all_first = uint32_t(tl.srcs.size());
                    for (size_t a = 0; a < n.children.size(); ++a) {
                        const int32_t c = n.children[a];
                        const PNode& cn = P->nodes[c];
                        TaskChild& tc = t.child[a];
                        tc.buf = -1;
                        if (cn.kind == PK_FOLD) {
                            tc.kind = CK_FOLD;
                            tc.coef = (trav == cn.last_to_act) ? -float(cn.value) : float(cn.value);  // cfr.rs:525-531
                        } else if (cn.kind == PK_SHOWDOWN) {
                            tc.kind = CK_SHOWDOWN;
                            tc.coef = float(cn.value);  // cfr.rs:532-543
                        } else {
                            std::vector<TaskSrc> ss;
                            if (!collect(c, ss)) return false;
                            if (ss.size() > 255 || tl.srcs.size() + ss.size() > 65535) {
                                err = "too many value sources for one task";
                                return false;
                            }
                            tc.kind = CK_VALUE;
                            tc.src_first = uint16_t(tl.srcs.size());
                            tc.n_src = uint8_t(ss.size());
                            tl.srcs.insert(tl.srcs.end(), ss.begin(), ss.end());
                        }
                    }
                    t.n_src_all = uint16_t(tl.srcs.size() - t.src_all_first);
                }
                for (uint32_t s = 0; s < P->segs[k].size(); ++s)
                    if (k > 0 && P->segs[k][s].root == pe.pnode) t.root_scatter = 1;
                const uint32_t ti = emit(t);
                up_task[pe.pnode] = int32_t(ti);
                for (uint32_t s = 0; s < P->segs[k].size(); ++s)
                    if (P->segs[k][s].root == pe.pnode) segroot_task[k][s] = int32_t(ti);
            }
            if (k > 0) {
                // gather the roots of street k into the chance leaves of street k-1 (cfr.rs:502-522)
                for (uint32_t l = 0; l < P->segs[k].size(); ++l) {
                    NodeTask t = blank(TK_GATHER, uint32_t(k - 1));
                    t.out = int32_t(l);
                    t.aux = segroot_cbuf[k][l];
                    if (!add_dep(t, segroot_task[k][l], DK_CHILD_BOARDS)) return false;
                    leaf_gather[k - 1][l] = int32_t(emit(t));
                }
                if (P->world > 1 && uint32_t(k) == P->shard_round) tl.phase_cut = tl.n_tickets;
            }
        }
        if (!(P->world > 1 && P->shard_round >= 1)) tl.phase_cut = tl.n_tickets;
        tl.root_cbuf = segroot_cbuf[0][0];
        // NOTE: dep fields hold node-task INDICES here; the engine converts them to first tickets when it
        // materialises the ticket numbering for a given instance count per round (engine.cu: materialize)
        if (tl.max_terminal > MAX_TERMINAL_CHILDREN) {
            err = "more than 3 terminal children under one action node";
            return false;
        }
        build_street();
        return true;
    }

    // ---- the final round as fused street programs (street.h) -------------------------------------
    struct StreetSeg {  // one segment while it is being compiled; rows and slots are local to the segment
        std::vector<SwDown> downs;
        std::vector<SwUp> ups;
        std::vector<SwTerm> terms;
        std::vector<uint8_t> row_need_y, row_need_m;
        uint32_t n_slots = 0;
        int32_t root_row = 0;
    };

    int16_t street_row(StreetSeg& g) {
        g.row_need_y.push_back(0);
        g.row_need_m.push_back(0);
        return int16_t(g.row_need_y.size() - 1);
    }
    float fold_coef(const PNode& cn) const { return (trav == cn.last_to_act) ? -float(cn.value) : float(cn.value); }  // cfr.rs:525-531

    // terms that add up to the value of opponent node `id` (sigma is inside the child rows, cfr.rs:583-588)
    void street_opp(StreetSeg& g, int32_t id, int16_t row, std::vector<SwTerm>& value_terms) {
        const PNode& n = P->nodes[id];
        SwDown d;
        std::memset(&d, 0, sizeof(d));
        d.n_act = uint8_t(n.children.size());
        d.in_row = row;
        d.cum_a = n.cum_a;
        for (int a = 0; a < SW_MAX_ACT; ++a) d.out_row[a] = -1;
        for (size_t a = 0; a < n.children.size(); ++a) d.out_row[a] = street_row(g);
        const size_t di = g.downs.size();
        g.downs.push_back(d);  // pre-order: the rows are written before any op below reads them
        (void)di;
        for (size_t a = 0; a < n.children.size(); ++a) {
            const int32_t c = n.children[a];
            const PNode& cn = P->nodes[c];
            const int16_t rc = d.out_row[a];
            if (cn.kind == PK_FOLD) {
                g.row_need_m[rc] = 1;
                value_terms.push_back(SwTerm{ST_FOLD, 0, rc, fold_coef(cn)});
            } else if (cn.kind == PK_SHOWDOWN) {
                g.row_need_y[rc] = 1;
                value_terms.push_back(SwTerm{ST_SHOWDOWN, 0, rc, float(cn.value)});  // cfr.rs:532-543
            } else {
                const int16_t slot = street_trav(g, c, rc, false);
                value_terms.push_back(SwTerm{ST_VALUE, 0, slot, 0.f});
            }
        }
    }

    // traverser node: returns the value slot it writes (-1 for the segment root)
    int16_t street_trav(StreetSeg& g, int32_t id, int16_t row, bool is_root) {
        const PNode& n = P->nodes[id];
        g.row_need_m[row] = 1;  // the strategy-sum weight is the compatible opponent reach (cfr.rs:618-619)
        std::vector<std::vector<SwTerm>> per_action(n.children.size());
        for (size_t a = 0; a < n.children.size(); ++a) {
            const int32_t c = n.children[a];
            const PNode& cn = P->nodes[c];
            if (cn.kind == PK_FOLD) {
                per_action[a].push_back(SwTerm{ST_FOLD, 0, row, fold_coef(cn)});
            } else if (cn.kind == PK_SHOWDOWN) {
                g.row_need_y[row] = 1;
                per_action[a].push_back(SwTerm{ST_SHOWDOWN, 0, row, float(cn.value)});
            } else {
                street_opp(g, c, row, per_action[a]);  // the opponent node sees the same reach
            }
        }
        SwUp u;
        std::memset(&u, 0, sizeof(u));
        u.kind = SU_TRAV;
        u.n_act = uint8_t(n.children.size());
        u.own_row = row;
        u.cum_a = n.cum_a;
        u.out_slot = is_root ? int16_t(-1) : int16_t(g.n_slots++);
        for (size_t a = 0; a < n.children.size(); ++a) {
            u.term_first[a] = uint16_t(g.terms.size());
            g.terms.insert(g.terms.end(), per_action[a].begin(), per_action[a].end());
        }
        for (size_t a = n.children.size(); a <= size_t(SW_MAX_ACT); ++a) u.term_first[a] = uint16_t(g.terms.size());
        g.ups.push_back(u);  // post-order: every slot it reads was written by an earlier op
        return u.out_slot;
    }

    // One list walk as program words (street.h).  classes: ascending strength; adds = opponent positions of the class
    // inside this list, emits = emit indices of the traverser hands of the class inside this list.
    struct ProgClass {
        std::vector<uint32_t> adds, emits;
    };
    static void schedule_program(const std::vector<ProgClass>& classes, uint32_t zero_pos, uint32_t dump_idx, std::vector<uint32_t>& words) {
        words.clear();
        std::vector<uint32_t> pending;  // emits of the last emitting class that have not found a step yet
        size_t pend_at = 0;
        auto pop = [&]() -> uint32_t {
            if (pend_at < pending.size()) return pending[pend_at++];
            return dump_idx;
        };
        auto left = [&]() { return pending.size() - pend_at; };
        auto step = [&](uint32_t add, uint32_t emit, uint32_t fl) { words.push_back(add | (emit << SW_EMIT_SHIFT) | fl); };
        for (const ProgClass& c : classes) {
            if (c.emits.empty()) {
                for (uint32_t a : c.adds) step(a, pop(), 0);
                continue;
            }
            const size_t n = std::max<size_t>(1, c.adds.size());
            // the step that ends this class overwrites m: emits of the class before must be out by then
            while (left() > n - 1) step(zero_pos, pop(), 0);
            for (size_t i = 0; i < n; ++i) {
                const uint32_t a = i < c.adds.size() ? c.adds[i] : zero_pos;
                const uint32_t fl = (i == 0 ? SW_CLASS_START : 0u) | (i == n - 1 ? SW_CLASS_END : 0u);
                step(a, i == n - 1 ? c.emits[0] : pop(), fl);
            }
            pending.assign(c.emits.begin() + 1, c.emits.end());
            pend_at = 0;
        }
        while (left() > 0) step(zero_pos, pop(), 0);
    }

    void build_street() {
        StreetPlan& S = sp;
        const uint32_t R = P->n_rounds, k = R - 1;
        S = StreetPlan();
        S.round_k = k;
        auto no = [&](const char* w) {
            S.eligible = false;
            S.why = w;
        };
        if (!(P->flags & RS_FLAG_STREET_KERNEL) && !getenv("RS_STREET")) return no("RS_FLAG_STREET_KERNEL is not set");
        if (P->sd[0].n_live.empty() || P->sd[1].n_live.empty()) return no("the final round has no showdown order");
        if (!P->loc[k][0].identity || !P->loc[k][1].identity) return no("the final round's tables are bucketed");
        for (const Segment& sg : P->segs[k]) {
            std::vector<int32_t> st{sg.root};
            while (!st.empty()) {
                const PNode& n = P->nodes[st.back()];
                st.pop_back();
                if (n.kind == PK_CHANCE) return no("chance node inside the final round");
                if (n.kind == PK_ACTION && n.children.size() > size_t(SW_MAX_ACT)) return no("a final-round node has more than 5 actions");
                for (int32_t c : n.children) st.push_back(c);
            }
        }
        // compile every segment on its own
        for (uint32_t s = 0; s < P->segs[k].size(); ++s) {
            StreetSeg g;
            const int32_t rootid = P->segs[k][s].root;
            const PNode& rn = P->nodes[rootid];
            g.root_row = street_row(g);
            if (rn.kind == PK_ACTION && rn.player == trav) {
                street_trav(g, rootid, int16_t(g.root_row), true);
            } else {
                std::vector<SwTerm> vt;
                if (rn.kind == PK_ACTION) {
                    street_opp(g, rootid, int16_t(g.root_row), vt);
                } else if (rn.kind == PK_SHOWDOWN) {  // bare showdown behind an all-in run-out
                    g.row_need_y[g.root_row] = 1;
                    vt.push_back(SwTerm{ST_SHOWDOWN, 0, int16_t(g.root_row), float(rn.value)});
                } else {
                    return no("unexpected segment root in the final round");
                }
                SwUp u;
                std::memset(&u, 0, sizeof(u));
                u.kind = SU_SUM;
                u.n_act = 1;
                u.own_row = -1;
                u.out_slot = -1;
                u.term_first[0] = uint16_t(g.terms.size());
                g.terms.insert(g.terms.end(), vt.begin(), vt.end());
                for (int a = 1; a <= SW_MAX_ACT; ++a) u.term_first[a] = uint16_t(g.terms.size());
                g.ups.push_back(u);
            }
            // renumber the rows: showdown rows, mass-only rows (both padded to quads), then the rest
            const size_t nr = g.row_need_y.size();
            std::vector<int16_t> newid(nr, -1);
            uint32_t n_sd = 0, n_mo = 0, sd_need_m = 0;
            for (size_t r = 0; r < nr; ++r)
                if (g.row_need_y[r]) {
                    if (g.row_need_m[r]) sd_need_m |= 1u << n_sd;
                    newid[r] = int16_t(n_sd++);
                }
            if (n_sd > uint32_t(SW_MAX_SD_ROWS)) return no("a final-round segment has more than 32 showdown rows");
            const uint32_t nq_sd = (n_sd + 3) / 4;
            for (size_t r = 0; r < nr; ++r)
                if (!g.row_need_y[r] && g.row_need_m[r]) newid[r] = int16_t(4 * nq_sd + n_mo++);
            const uint32_t nq_mo = (n_mo + 3) / 4;
            uint32_t next = 4 * (nq_sd + nq_mo);
            for (size_t r = 0; r < nr; ++r)
                if (newid[r] < 0) newid[r] = int16_t(next++);
            if (next > 2000) return no("a final-round segment needs too many reach rows");
            if (S.terms.size() + g.terms.size() > 65000) return no("too many value terms");
            SwSeg seg;
            std::memset(&seg, 0, sizeof(seg));
            seg.down_first = uint32_t(S.downs.size());
            seg.down_count = uint32_t(g.downs.size());
            seg.up_first = uint32_t(S.ups.size());
            seg.up_count = uint32_t(g.ups.size());
            seg.root_row = newid[g.root_row];
            seg.root_in = k > 0 ? leaf_rbuf[k - 1][s] : RIN_INITIAL;
            seg.root_out = segroot_cbuf[k][s];
            seg.n_rows = next;
            seg.nq_sd = nq_sd;
            seg.nq_mo = nq_mo;
            seg.n_slots = g.n_slots;
            seg.sd_need_m = sd_need_m;
            const uint16_t term0 = uint16_t(S.terms.size());
            for (SwDown d : g.downs) {
                d.in_row = newid[d.in_row];
                for (int a = 0; a < d.n_act; ++a) d.out_row[a] = newid[d.out_row[a]];
                S.downs.push_back(d);
            }
            for (SwUp u : g.ups) {
                if (u.own_row >= 0) u.own_row = newid[u.own_row];
                for (int a = 0; a <= SW_MAX_ACT; ++a) u.term_first[a] = uint16_t(u.term_first[a] + term0);
                S.ups.push_back(u);
            }
            for (SwTerm t : g.terms) {
                if (t.kind != ST_VALUE) t.id = newid[t.id];
                S.terms.push_back(t);
            }
            S.segs.push_back(seg);
            S.max_rows = std::max(S.max_rows, seg.n_rows);
            S.max_slots = std::max(S.max_slots, seg.n_slots);
            S.max_q_sd = std::max(S.max_q_sd, nq_sd);
            S.max_q_mo = std::max(S.max_q_mo, nq_mo);
        }
        // list programs of every local board: the 52 card lists in SW_LIST_PIECES pieces each, then the global strength
        // order in SW_CHUNKS pieces
        const int o = 1 - trav;
        const LocalTables& Lp = P->loc[k][trav];
        const LocalTables& Lo = P->loc[k][o];
        const uint32_t Hp = P->H[trav], Ho = P->H[o];
        const uint32_t HpP = Lp.Hpad, HoP = Lo.Hpad;
        if (HoP >= SW_ADD_MASK || 2 * (HpP + 1) > SW_EMIT_MASK) return no("range too large for the program words");
        const uint32_t lo_b = P->local_lo[k], hi_b = P->local_hi[k];
        S.prog_off.assign(size_t(hi_b - lo_b) + 1, 0);
        S.l_steps.assign(hi_b - lo_b, 0);
        S.c_steps.assign(hi_b - lo_b, 0);
        S.hinfo.assign(size_t(hi_b - lo_b) * HpP * 2, 0);
        const uint32_t zero_pos = HoP, dump_idx = HpP;
        std::vector<ProgClass> cls;
        constexpr int NLP = SW_CARDS * SW_LIST_PIECES;
        std::vector<std::vector<uint32_t>> lw(NLP), cw(SW_CHUNKS);
        // Cut the classes of one list into at most n_pieces pieces of about equal length and schedule each piece on its own.
        // A class far longer than a piece (a board that plays: hundreds of hands tie) becomes a RUN of pieces that only
        // add and combine (m stays 0): its hands take A + B = base[first piece of the run] + base[piece after the run].
        // on_hand(emit index, lo piece, hi piece): normal classes lo == hi == their piece.
        auto cut_list = [&](std::vector<ProgClass>& classes, uint32_t n_pieces, std::vector<uint32_t>* words, uint32_t& max_steps,
                            const std::function<void(uint32_t, uint32_t, uint32_t)>& on_hand) {
            // steps a class costs: its adds, and its emits but the first trail behind on steps of their own when the classes
            // after it are short (schedule_program)
            auto len_of = [](const ProgClass& c) { return uint64_t(std::max<size_t>(1, c.adds.size() + (c.emits.empty() ? 0 : c.emits.size() - 1))); };
            // pieces used when no piece may be longer than `cap`: classes are packed greedily, a class longer than cap
            // becomes a run of ceil(len / cap) pieces
            auto pieces_needed = [&](uint64_t cap) {
                uint64_t n = 0, acc = 0;
                for (const ProgClass& c : classes) {
                    const uint64_t len = len_of(c);
                    if (len > cap) {
                        if (acc) ++n, acc = 0;
                        n += (len + cap - 1) / cap;
                    } else {
                        if (acc + len > cap) ++n, acc = 0;
                        acc += len;
                    }
                }
                return n + (acc ? 1 : 0);
            };
            uint64_t lo = 1, hi = 1;
            for (const ProgClass& c : classes) hi += len_of(c);
            while (lo < hi) {  // smallest cap that fits n_pieces
                const uint64_t mid = (lo + hi) / 2;
                if (pieces_needed(mid) <= n_pieces) hi = mid;
                else lo = mid + 1;
            }
            const uint64_t cap = lo;
            uint32_t ch = 0;
            std::vector<ProgClass> piece;
            uint64_t acc = 0;
            auto flush = [&]() {
                if (piece.empty()) return;
                schedule_program(piece, zero_pos, dump_idx, words[ch]);
                max_steps = std::max<uint32_t>(max_steps, uint32_t(words[ch].size()));
                for (const ProgClass& c : piece)
                    for (uint32_t e : c.emits) on_hand(e, ch, ch);
                piece.clear();
                acc = 0;
                ++ch;
            };
            for (size_t ci = 0; ci < classes.size(); ++ci) {
                ProgClass& c = classes[ci];
                const uint64_t len = len_of(c);
                if (len > cap) {
                    flush();
                    const uint32_t r = uint32_t((len + cap - 1) / cap);
                    const uint32_t c1 = ch;
                    for (uint32_t t = 0; t < r; ++t) {
                        const size_t a0 = c.adds.size() * t / r, a1 = c.adds.size() * (t + 1) / r;
                        const size_t e0 = c.emits.size() * t / r, e1 = c.emits.size() * (t + 1) / r;
                        std::vector<uint32_t>& wv = words[ch];
                        wv.clear();
                        for (size_t i = 0; i < std::max(a1 - a0, e1 - e0); ++i)
                            wv.push_back((a0 + i < a1 ? c.adds[a0 + i] : zero_pos) | ((e0 + i < e1 ? c.emits[e0 + i] : dump_idx) << SW_EMIT_SHIFT));
                        max_steps = std::max<uint32_t>(max_steps, uint32_t(wv.size()));
                        ++ch;
                    }
                    for (uint32_t e : c.emits) on_hand(e, c1, ch);
                    continue;
                }
                if (acc + len > cap) flush();
                acc += len;
                piece.push_back(std::move(c));
            }
            flush();
            for (uint32_t c2 = ch; c2 < n_pieces; ++c2) words[c2].clear();
        };
        for (uint32_t b = lo_b; b < hi_b; ++b) {
            const uint32_t np = Lp.n_live[b], no_ = Lo.n_live[b];
            const uint16_t* sp_ = &Lp.slot_of_pos[size_t(b) * Lp.Hpad];
            const uint16_t* so_ = &Lo.slot_of_pos[size_t(b) * Lo.Hpad];
            const uint32_t* strp = &P->sd[trav].strength[size_t(b) * Hp];
            const uint32_t* stro = &P->sd[o].strength[size_t(b) * Ho];
            // positions ascend with strength on the final round (LocalTables): a merge of the two orders gives the classes
            auto classes_of = [&](int card, std::vector<ProgClass>& out) {
                out.clear();
                uint32_t i = 0, j = 0;
                auto skip_p = [&]() {
                    while (card >= 0 && i < np && P->hand_cards[trav][2 * sp_[i]] != card && P->hand_cards[trav][2 * sp_[i] + 1] != card) ++i;
                };
                auto skip_o = [&]() {
                    while (card >= 0 && j < no_ && P->hand_cards[o][2 * so_[j]] != card && P->hand_cards[o][2 * so_[j] + 1] != card) ++j;
                };
                skip_p();
                skip_o();
                while (i < np || j < no_) {
                    uint32_t st = 0xffffffffu;
                    if (i < np) st = std::min(st, strp[sp_[i]]);
                    if (j < no_) st = std::min(st, stro[so_[j]]);
                    ProgClass c;
                    while (j < no_ && stro[so_[j]] == st) {
                        c.adds.push_back(j);
                        ++j;
                        skip_o();
                    }
                    while (i < np && strp[sp_[i]] == st) {
                        const uint32_t tgt = (card >= 0 && P->hand_cards[trav][2 * sp_[i] + 1] == card) ? 1u : 0u;
                        c.emits.push_back(tgt * (HpP + 1) + i);
                        ++i;
                        skip_p();
                    }
                    out.push_back(std::move(c));
                }
            };
            uint32_t* hi_ = &S.hinfo[size_t(b - lo_b) * HpP * 2];
            uint32_t ls = 0, cs = 0;
            for (int c = 0; c < SW_CARDS; ++c) {
                classes_of(c, cls);
                cut_list(cls, SW_LIST_PIECES, &lw[size_t(c) * SW_LIST_PIECES], ls, [&](uint32_t e, uint32_t lo, uint32_t hi) {
                    const bool second = e > HpP;  // emit index of target 1: the list is the hand's second card
                    const uint32_t pos = second ? e - (HpP + 1) : e;
                    hi_[2 * pos] |= second ? (lo << SW_HI_P1LO_SHIFT) | (hi << SW_HI_P1HI_SHIFT) : (lo << SW_HI_P0LO_SHIFT) | (hi << SW_HI_P0HI_SHIFT);
                });
            }
            classes_of(-1, cls);
            cut_list(cls, SW_CHUNKS, cw.data(), cs, [&](uint32_t e, uint32_t lo, uint32_t hi) { hi_[2 * e + 1] |= lo | (hi << SW_HI_CHHI_SHIFT); });
            for (uint32_t i = 0; i < np; ++i) {
                const uint32_t h = sp_[i];
                uint32_t same = HoP;
                const uint16_t sm = P->same[trav][h];
                if (sm != 0xFFFF) {
                    const uint16_t op = Lo.pos_of_slot[size_t(b) * Ho + sm];
                    if (op < no_) same = op;
                }
                hi_[2 * i] |= uint32_t(P->hand_cards[trav][2 * h]) | (uint32_t(P->hand_cards[trav][2 * h + 1]) << SW_HI_C1_SHIFT);
                hi_[2 * i + 1] |= same << SW_HI_SAME_SHIFT;
            }
            for (uint32_t i = np; i < HpP; ++i) hi_[2 * i + 1] = HoP << SW_HI_SAME_SHIFT;
            ls = (ls + 3) & ~3u;  // the walks fetch four words at a time
            cs = (cs + 3) & ~3u;
            // pack: [ls][208] then [cs][128]; short programs are padded with steps that add nothing and emit to the dump cell
            const uint32_t nop = zero_pos | (dump_idx << SW_EMIT_SHIFT);
            const size_t base = S.prog.size();
            S.prog.resize(base + size_t(ls) * NLP + size_t(cs) * SW_CHUNKS, nop);
            for (int c = 0; c < NLP; ++c)
                for (size_t t = 0; t < lw[c].size(); ++t) S.prog[base + t * NLP + c] = lw[c][t];
            const size_t cbase = base + size_t(ls) * NLP;
            for (int c = 0; c < SW_CHUNKS; ++c)
                for (size_t t = 0; t < cw[c].size(); ++t) S.prog[cbase + t * SW_CHUNKS + c] = cw[c][t];
            if (S.prog.size() > 0xfffffff0ull) return no("list programs exceed 32-bit word offsets");
            S.prog_off[b - lo_b] = uint32_t(base);
            S.l_steps[b - lo_b] = ls;
            S.c_steps[b - lo_b] = cs;
        }
        S.prog_off[hi_b - lo_b] = uint32_t(S.prog.size());
        S.eligible = true;
    }
};

}  // namespace

bool compile_plan(const rs_tree* tree, const rs_ranges* ranges, const rs_abstraction* abs,
                  const rs_config* cfg, const uint64_t* board_masks, uint32_t n_sub, Plan* P,
                  std::string* err, BatchIndexFn batch_index) {
    auto fail = [&](const std::string& m) {
        if (err) *err = m;
        return false;
    };
    if (!tree || !ranges || !cfg || !P) return fail("null argument");
    if (!tree->type || !tree->parent || !tree->child_offset || !tree->children || !tree->player ||
        !tree->an_index || !tree->round_idx || !tree->value || !tree->ttype || !tree->last_to_act)
        return fail("rs_tree has a null array");
    if (tree->n_nodes < 2) return fail("tree needs at least a root and one child");
    if (n_sub == 0 || !board_masks) return fail("need at least one board");
    int nb0 = __builtin_popcountll(board_masks[0]);
    if (nb0 < 3 || nb0 > 5) return fail("invalid board mask");  // state.rs:63
    for (uint32_t i = 0; i < n_sub; ++i) {
        if (__builtin_popcountll(board_masks[i]) != nb0) return fail("all subgame boards must have the same number of cards");
        if (board_masks[i] >> 52) return fail("board mask has bits above card 51");
    }
    *P = Plan();
    P->first_round = uint32_t(nb0 - 3);
    P->n_sub = n_sub;
    P->flags = cfg->flags;
    P->rank = cfg->world_size > 1 ? cfg->rank : 0;
    P->world = cfg->world_size > 1 ? cfg->world_size : 1;
    if (P->rank < 0 || P->rank >= P->world) return fail("rank out of range");

    // ---- tree ----
    if (tree->type[0] != RS_NODE_PRIVATE_CHANCE) return fail("root must be the private chance node (tree_builder.rs:60-66)");
    if (tree->child_offset[1] - tree->child_offset[0] != 1) return fail("private chance root needs exactly one child");
    Builder B{tree, P};
    int32_t root = B.copy_node(tree->children[tree->child_offset[0]], 0, 0);
    if (root < 0) return fail(B.err);
    P->n_rounds = B.max_round + 1;
    if (P->first_round + P->n_rounds > 3) return fail("tree has more betting rounds than the board allows");
    B.make_segment(0, root);

    // action-node numbering per (round, player), in ActionNode.index order
    {
        uint32_t max_an = 0;
        for (auto& n : P->nodes)
            if (n.kind == PK_ACTION) max_an = std::max(max_an, n.an_index + 1);
        P->an_to_pnode.assign(max_an, -1);
        for (size_t i = 0; i < P->nodes.size(); ++i)
            if (P->nodes[i].kind == PK_ACTION) {
                if (P->an_to_pnode[P->nodes[i].an_index] != -1) return fail("duplicate ActionNode.index");
                P->an_to_pnode[P->nodes[i].an_index] = int32_t(i);
            }
        for (uint32_t an = 0; an < max_an; ++an) {
            int32_t id = P->an_to_pnode[an];
            if (id < 0) continue;
            PNode& n = P->nodes[id];
            if (n.children.size() > 8) return fail("more than 8 actions at a node is not supported");
            RoundPlayerTables& T = P->tabs[n.round_k][n.player];
            n.tab_j = int32_t(T.n_nodes++);
            n.cum_a = T.sum_a;
            T.sum_a += uint32_t(n.children.size());
            T.node_an_index.push_back(an);
            T.node_n_act.push_back(uint32_t(n.children.size()));
        }
    }

    // ---- hands ----
    for (int q = 0; q < 2; ++q) {
        uint32_t H = ranges->n_hands[q];
        if (H == 0 || H > 1326) return fail("range size must be in 1..1326");
        if (!ranges->hands[q]) return fail("null range");
        P->H[q] = H;
        P->hand_cards[q].assign(ranges->hands[q], ranges->hands[q] + 2 * size_t(H));
        std::vector<uint8_t> seen(52 * 52, 0);
        for (uint32_t h = 0; h < H; ++h) {
            uint8_t a = P->hand_cards[q][2 * h], b = P->hand_cards[q][2 * h + 1];
            if (a >= 52 || b >= 52 || a == b) return fail("bad hole cards in range");
            int hi = std::max(a, b), lo = std::min(a, b);
            if (seen[hi * 52 + lo]) return fail("duplicate combo in range");
            seen[hi * 52 + lo] = 1;
        }
    }
    for (int q = 0; q < 2; ++q) {
        std::vector<int32_t> slot(52 * 52, -1);
        const auto& oc = P->hand_cards[1 - q];
        for (uint32_t h = 0; h < P->H[1 - q]; ++h) {
            int a = oc[2 * h], b = oc[2 * h + 1];
            slot[std::max(a, b) * 52 + std::min(a, b)] = int32_t(h);
        }
        P->same[q].assign(P->H[q], 0xFFFF);
        P->card_hands[q].assign(52 * 52, 0xFFFF);
        uint32_t cnt[52] = {0};
        for (uint32_t h = 0; h < P->H[q]; ++h) {
            int a = P->hand_cards[q][2 * h], b = P->hand_cards[q][2 * h + 1];
            int32_t s = slot[std::max(a, b) * 52 + std::min(a, b)];
            if (s >= 0) P->same[q][h] = uint16_t(s);
            P->card_hands[q][a * 52 + cnt[a]++] = uint16_t(h);
            P->card_hands[q][b * 52 + cnt[b]++] = uint16_t(h);
        }
    }

    // ---- boards (board_table, README.md:41-43) ----
    P->n_boards[0] = n_sub;
    P->board_mask[0].assign(board_masks, board_masks + n_sub);
    P->board_parent[0].assign(n_sub, -1);
    P->board_card[0].assign(n_sub, 0xFF);
    for (uint32_t k = 1; k < P->n_rounds; ++k) {
        uint32_t per = uint32_t(52 - (nb0 + int(k) - 1));
        P->deal_count[k] = per;
        P->n_boards[k] = P->n_boards[k - 1] * per;
        P->board_mask[k].reserve(P->n_boards[k]);
        for (uint32_t pb = 0; pb < P->n_boards[k - 1]; ++pb) {
            uint64_t pm = P->board_mask[k - 1][pb];
            for (int c = 0; c < 52; ++c) {  // ascending card order (cfr.rs:63-68)
                if (pm & (1ull << c)) continue;
                P->board_mask[k].push_back(pm | (1ull << c));
                P->board_parent[k].push_back(int32_t(pb));
                P->board_card[k].push_back(uint8_t(c));
            }
        }
    }
    // sharding
    for (uint32_t k = 0; k < P->n_rounds; ++k) {
        P->local_lo[k] = 0;
        P->local_hi[k] = P->n_boards[k];
    }
    if (P->world > 1) {
        if (n_sub > 1) {
            P->shard_round = 0;
        } else {
            if (P->n_rounds < 2) return fail("a single-board river subgame has nothing to shard across GPUs");
            P->shard_round = 1;
        }
        uint32_t sr = P->shard_round;
        uint64_t n = P->n_boards[sr];
        if (n < uint64_t(P->world)) return fail("fewer boards than ranks at the sharded level");
        P->local_lo[sr] = uint32_t(n * uint64_t(P->rank) / uint64_t(P->world));
        P->local_hi[sr] = uint32_t(n * uint64_t(P->rank + 1) / uint64_t(P->world));
        for (uint32_t k = sr + 1; k < P->n_rounds; ++k) {
            P->local_lo[k] = P->local_lo[k - 1] * P->deal_count[k];
            P->local_hi[k] = P->local_hi[k - 1] * P->deal_count[k];
        }
    }

    // ---- chance weights (cfr.rs:491, 510) ----
    P->n_combos.assign(n_sub, 0);
    P->chance_scale[0].assign(n_sub, 0.f);
    for (uint32_t s = 0; s < n_sub; ++s) {
        uint64_t bm = board_masks[s];
        // generate_all_hole_card_combos (cfr.rs:73-98): non-overlapping (h0,h1) pairs.
        // count = live0*live1 - sum_c n0[c]*n1[c] + shared identical combos
        uint64_t live[2] = {0, 0};
        uint64_t nc[2][52] = {{0}};
        for (int q = 0; q < 2; ++q)
            for (uint32_t h = 0; h < P->H[q]; ++h) {
                int a = P->hand_cards[q][2 * h], b = P->hand_cards[q][2 * h + 1];
                if (bm & ((1ull << a) | (1ull << b))) continue;
                live[q]++;
                nc[q][a]++;
                nc[q][b]++;
            }
        uint64_t overlap = 0;
        for (int c = 0; c < 52; ++c) overlap += nc[0][c] * nc[1][c];
        uint64_t ident = 0;
        for (uint32_t h = 0; h < P->H[0]; ++h) {
            int a = P->hand_cards[0][2 * h], b = P->hand_cards[0][2 * h + 1];
            if (bm & ((1ull << a) | (1ull << b))) continue;
            if (P->same[0][h] != 0xFFFF) ident++;
        }
        uint64_t n = live[0] * live[1] - overlap + ident;
        if (n == 0) return fail("no compatible hole-card combos for a subgame");
        P->n_combos[s] = n;
        P->chance_scale[0][s] = float(1.0 / double(n));
    }
    for (uint32_t k = 1; k < P->n_rounds; ++k) {
        P->chance_scale[k].resize(P->n_boards[k]);
        double len = double(52 - (nb0 + int(k) - 1) - 4);  // cfr.rs:49-70: 52 - board - both hands
        for (uint32_t b = 0; b < P->n_boards[k]; ++b)
            P->chance_scale[k][b] = float(double(P->chance_scale[k - 1][P->board_parent[k][b]]) / len);
    }
    // NOTE: chance_scale[k] for k>=1 is derived in double from the fp32 parent to stay
    // identical on every rank; the oracle uses the same fp32 constants.

    // ---- card tables (README.md:36-39; card_abstraction.rs:75-184, 204-209) ----
    for (uint32_t k = 0; k < P->n_rounds; ++k) {
        uint32_t kind = RS_ABS_NONE;
        const rs_round_abstraction* ra = nullptr;
        if (abs && k < abs->n_rounds) {
            ra = &abs->rounds[k];
            kind = ra->kind;
        }
        if (kind > RS_ABS_BUCKET_TABLE) return fail("unknown abstraction kind");
        HandIndexer indexer;
        int nbc = nb0 + int(k);
        if (kind == RS_ABS_ISOMORPHIC || kind == RS_ABS_CLUSTER_ARR) {
            // hand_indexer_s::init(2, [2, 3|4|5]) (card_abstraction.rs:88-90)
            if (!indexer.init(2, {2, uint8_t(nbc)})) return fail("hand indexer init failed");
            if (kind == RS_ABS_CLUSTER_ARR) {
                if (!ra->cluster_arr) return fail("cluster_arr is null");
                if (ra->cluster_arr_len < indexer.size(1)) return fail("cluster_arr shorter than the round's canonical-hand count");
            }
        }
        for (int q = 0; q < 2; ++q) {
            RoundPlayerTables& T = P->tabs[k][q];
            uint32_t H = P->H[q];
            uint32_t nB = P->n_boards[k];
            if (kind == RS_ABS_BUCKET_TABLE && !ra->bucket_table[q]) return fail("bucket_table is null");
            T.row_of_hand.assign(size_t(nB) * H, 0xFFFF);
            T.row_start.assign(size_t(nB) * (H + 1), 0);
            T.row_hands.assign(size_t(nB) * H, 0xFFFF);
            T.n_rows.assign(nB, 0);
            T.board_off.assign(size_t(nB) + 1, 0);
            // canonical hand indices of every live (board, hand) of this player, in one batch: the device indexer
            // when the engine provides it (indexer_kernel.cu), the host indexer otherwise
            std::vector<uint64_t> canon;  // [(b - local_lo) * H + h]
            if (kind == RS_ABS_ISOMORPHIC || kind == RS_ABS_CLUSTER_ARR) {
                const uint32_t lo = P->local_lo[k], hi = P->local_hi[k];
                const int nc = 2 + nbc;
                std::vector<uint8_t> cards_all;
                std::vector<uint32_t> where;
                cards_all.reserve(size_t(hi - lo) * H * nc);
                where.reserve(size_t(hi - lo) * H);
                for (uint32_t b = lo; b < hi; ++b) {
                    const uint64_t bm = P->board_mask[k][b];
                    uint8_t bc[5];
                    int nbd = 0;
                    for (uint64_t m = bm; m; m &= m - 1) bc[nbd++] = uint8_t(__builtin_ctzll(m));  // ascending (cfr.rs:78-82)
                    for (uint32_t h = 0; h < H; ++h) {
                        const uint8_t a = P->hand_cards[q][2 * h], c = P->hand_cards[q][2 * h + 1];
                        if (bm & ((1ull << a) | (1ull << c))) continue;
                        cards_all.push_back(a);
                        cards_all.push_back(c);
                        for (int i = 0; i < nbd; ++i) cards_all.push_back(bc[i]);
                        where.push_back((b - lo) * H + h);
                    }
                }
                std::vector<uint64_t> idx(where.size());
                if (batch_index) {
                    std::string berr;
                    if (!batch_index(indexer, cards_all.data(), where.size(), idx.data(), &berr)) return fail("device indexer: " + berr);
                } else {
                    for (size_t i = 0; i < where.size(); ++i) idx[i] = indexer.get_index(&cards_all[i * nc]);
                }
                canon.assign(size_t(hi - lo) * H, 0);
                for (size_t i = 0; i < where.size(); ++i) canon[where[i]] = idx[i];
            }
            std::unordered_map<uint64_t, uint32_t> dense;
            std::vector<uint32_t> row_count;
            for (uint32_t b = P->local_lo[k]; b < P->local_hi[k]; ++b) {
                uint64_t bm = P->board_mask[k][b];
                dense.clear();
                row_count.clear();
                uint16_t* roh = &T.row_of_hand[size_t(b) * H];
                for (uint32_t h = 0; h < H; ++h) {
                    uint8_t a = P->hand_cards[q][2 * h], c = P->hand_cards[q][2 * h + 1];
                    if (bm & ((1ull << a) | (1ull << c))) continue;
                    uint64_t key;
                    if (kind == RS_ABS_NONE) {
                        key = h;
                    } else if (kind == RS_ABS_BUCKET_TABLE) {
                        key = ra->bucket_table[q][size_t(b) * H + h];
                    } else {
                        key = canon[size_t(b - P->local_lo[k]) * H + h];  // hand_indexer.get_index (card_abstraction.rs:205)
                        if (kind == RS_ABS_CLUSTER_ARR) key = ra->cluster_arr[key];  // index_to_cluster, card_abstraction.rs:20-29
                    }
                    auto it = dense.find(key);
                    uint32_t row;
                    if (it == dense.end()) {  // first-seen order (deterministic here; channel order in the reference)
                        row = uint32_t(dense.size());
                        dense.emplace(key, row);
                        row_count.push_back(0);
                    } else {
                        row = it->second;
                    }
                    roh[h] = uint16_t(row);
                    row_count[row]++;
                }
                uint32_t nr = uint32_t(row_count.size());
                T.n_rows[b] = nr;
                uint16_t* rs_ = &T.row_start[size_t(b) * (H + 1)];
                uint32_t acc = 0;
                for (uint32_t r = 0; r < nr; ++r) {
                    rs_[r] = uint16_t(acc);
                    acc += row_count[r];
                }
                for (uint32_t r = nr; r <= H; ++r) rs_[r] = uint16_t(acc);
                std::vector<uint32_t> fill(nr, 0);
                uint16_t* rh = &T.row_hands[size_t(b) * H];
                for (uint32_t h = 0; h < H; ++h) {
                    uint16_t r = roh[h];
                    if (r == 0xFFFF) continue;
                    rh[rs_[r] + fill[r]++] = uint16_t(h);
                }
            }
        }
    }

    // ---- showdown order (final round must be the river) ----
    bool any_showdown = false;
    for (auto& n : P->nodes)
        if (n.kind == PK_SHOWDOWN) {
            any_showdown = true;
            if (P->first_round + n.round_k != RS_ROUND_RIVER) return fail("showdown before the river");
        }
    if (any_showdown) {
        const uint32_t k = P->n_rounds - 1;
        const uint32_t nB = P->n_boards[k];
        for (int q = 0; q < 2; ++q) {
            const uint32_t H = P->H[q];
            ShowdownTables& S = P->sd[q];
            S.sorted.assign(size_t(nB) * H, 0xFFFF);
            S.n_live.assign(nB, 0);
            S.cls.assign(size_t(nB) * H, 0);
            S.strength.assign(size_t(nB) * H, 0);
            std::vector<uint16_t> order;
            for (uint32_t b = P->local_lo[k]; b < P->local_hi[k]; ++b) {
                const uint64_t bm = P->board_mask[k][b];
                uint32_t* str = &S.strength[size_t(b) * H];
                order.clear();
                for (uint32_t h = 0; h < H; ++h) {
                    const uint64_t hm = (1ull << P->hand_cards[q][2 * h]) | (1ull << P->hand_cards[q][2 * h + 1]);
                    if (hm & bm) continue;
                    str[h] = evaluate_mask(bm | hm) + 1;  // TrainHand::get_hand + evaluate (cfr.rs:38-46, 534)
                    order.push_back(uint16_t(h));
                }
                std::stable_sort(order.begin(), order.end(), [&](uint16_t x, uint16_t y) { return str[x] < str[y]; });
                S.n_live[b] = uint32_t(order.size());
                uint32_t cls = 0;
                for (size_t i = 0; i < order.size(); ++i) {
                    if (i > 0 && str[order[i]] != str[order[i - 1]]) cls++;
                    S.sorted[size_t(b) * H + i] = order[i];
                    S.cls[size_t(b) * H + i] = cls;
                }
            }
        }
    }

    // ---- board-local hand order and the device tables indexed by it ----
    for (uint32_t k = 0; k < P->n_rounds; ++k) {
        const uint32_t nB = P->n_boards[k];
        const bool final_round = any_showdown && k == P->n_rounds - 1;
        for (int q = 0; q < 2; ++q) {
            LocalTables& L = P->loc[k][q];
            const RoundPlayerTables& T = P->tabs[k][q];
            const uint32_t H = P->H[q];
            L.Hpad = (H + 3) & ~3u;
            L.slot_of_pos.assign(size_t(nB) * L.Hpad, 0xFFFF);
            L.pos_of_slot.assign(size_t(nB) * H, 0xFFFF);
            L.n_live.assign(nB, 0);
            L.row_of_pos.assign(size_t(nB) * L.Hpad, 0xFFFF);
            L.row_start.assign(size_t(nB) * (L.Hpad + 4), 0);
            L.row_pos.assign(size_t(nB) * L.Hpad, 0xFFFF);
            L.n_rows_pad.assign(nB, 0);
            L.identity = true;
            for (uint32_t b = P->local_lo[k]; b < P->local_hi[k]; ++b) {
                uint16_t* sop = &L.slot_of_pos[size_t(b) * L.Hpad];
                uint16_t* pos = &L.pos_of_slot[size_t(b) * H];
                const uint16_t* roh = &T.row_of_hand[size_t(b) * H];
                uint32_t n = 0;
                if (final_round) {
                    const ShowdownTables& S = P->sd[q];
                    for (uint32_t i = 0; i < S.n_live[b]; ++i) sop[n++] = S.sorted[size_t(b) * H + i];
                } else {
                    for (uint32_t h = 0; h < H; ++h)
                        if (roh[h] != 0xFFFF) sop[n++] = uint16_t(h);
                }
                L.n_live[b] = n;
                for (uint32_t h = 0; h < H; ++h)
                    if (roh[h] == 0xFFFF) sop[n++] = uint16_t(h);  // hands the board removes go last
                for (uint32_t i = 0; i < H; ++i) pos[sop[i]] = uint16_t(i);
                // rows by position; NONE-abstraction tables are renumbered so that row == position
                uint16_t* rop = &L.row_of_pos[size_t(b) * L.Hpad];
                for (uint32_t i = 0; i < L.n_live[b]; ++i) rop[i] = roh[sop[i]];
            }
        }
    }
    // The lossless tables (one row per live hand) use the local position as the row id, which makes every
    // table access of a thread's four hands one contiguous run.  card_table (row_of_hand) reports that id.
    for (uint32_t k = 0; k < P->n_rounds; ++k) {
        uint32_t kind = (abs && k < abs->n_rounds) ? abs->rounds[k].kind : uint32_t(RS_ABS_NONE);
        for (int q = 0; q < 2; ++q) {
            LocalTables& L = P->loc[k][q];
            RoundPlayerTables& T = P->tabs[k][q];
            const uint32_t H = P->H[q];
            const uint32_t nB = P->n_boards[k];
            if (kind == RS_ABS_NONE) {
                for (uint32_t b = P->local_lo[k]; b < P->local_hi[k]; ++b) {
                    for (uint32_t i = 0; i < L.n_live[b]; ++i) {
                        L.row_of_pos[size_t(b) * L.Hpad + i] = uint16_t(i);
                        T.row_of_hand[size_t(b) * H + L.slot_of_pos[size_t(b) * L.Hpad + i]] = uint16_t(i);
                    }
                }
                L.identity = true;
            } else {
                L.identity = false;
            }
            uint64_t off = 0;
            for (uint32_t b = 0; b < nB; ++b) {
                T.board_off[b] = off;
                if (b < P->local_lo[k] || b >= P->local_hi[k]) continue;
                const uint32_t nr = T.n_rows[b];
                L.n_rows_pad[b] = (nr + 3) & ~3u;
                off += uint64_t(L.n_rows_pad[b]) * T.sum_a;  // [node][row_pad][A]: every slab starts 16-byte aligned
                // CSR row -> positions
                uint16_t* rs_ = &L.row_start[size_t(b) * (L.Hpad + 4)];
                uint16_t* rp = &L.row_pos[size_t(b) * L.Hpad];
                const uint16_t* rop = &L.row_of_pos[size_t(b) * L.Hpad];
                std::vector<uint32_t> cnt(nr + 1, 0);
                for (uint32_t i = 0; i < L.n_live[b]; ++i) cnt[rop[i] + 1]++;
                for (uint32_t r = 0; r < nr; ++r) cnt[r + 1] += cnt[r];
                for (uint32_t r = 0; r <= nr; ++r) rs_[r] = uint16_t(cnt[r]);
                for (uint32_t r = nr + 1; r < L.Hpad + 4; ++r) rs_[r] = uint16_t(cnt[nr]);
                std::vector<uint32_t> fill(cnt.begin(), cnt.end() - 1);
                for (uint32_t i = 0; i < L.n_live[b]; ++i) rp[fill[rop[i]]++] = uint16_t(i);
            }
            T.board_off[nB] = off;
        }
    }
    // per-card lists (opponent role), per-hand records (traverser role), street-transition maps
    for (uint32_t k = 0; k < P->n_rounds; ++k) {
        const uint32_t nB = P->n_boards[k];
        const bool final_round = any_showdown && k == P->n_rounds - 1;
        for (int q = 0; q < 2; ++q) {
            LocalTables& L = P->loc[k][q];
            L.cl_pos.assign(size_t(nB) * 2 * L.Hpad, uint16_t(CL_POS_NONE));
            L.hrec.assign(size_t(nB) * L.Hpad, HandRec{0, 0, 0, HREC_K_NONE | (HREC_K_NONE << 9) | (HREC_SAME_NONE << 18)});
            if (k > 0) {
                L.parent_pos.assign(size_t(nB) * L.Hpad, 0xFFFF);
                L.child_pos.assign(size_t(nB) * P->loc[k - 1][q].Hpad, 0xFFFF);
            }
        }
        for (uint32_t b = P->local_lo[k]; b < P->local_hi[k]; ++b) {
            uint32_t seg_start[2][53];
            int seg_ord[2][52];                    // ordinal of the card among the player's non-empty lists, -1 = empty
            std::vector<uint32_t> seg_str[2][52];  // strengths inside each card segment (final round)
            for (int q = 0; q < 2; ++q) {
                LocalTables& L = P->loc[k][q];
                const uint32_t H = P->H[q];
                const uint16_t* sop = &L.slot_of_pos[size_t(b) * L.Hpad];
                uint32_t cnt[53] = {0};
                for (uint32_t i = 0; i < L.n_live[b]; ++i) {
                    cnt[P->hand_cards[q][2 * sop[i]] + 1]++;
                    cnt[P->hand_cards[q][2 * sop[i] + 1] + 1]++;
                }
                seg_start[q][0] = 0;
                for (int c = 0; c < 52; ++c) seg_start[q][c + 1] = seg_start[q][c] + cnt[c + 1];
                uint32_t fill[52];
                for (int c = 0; c < 52; ++c) fill[c] = seg_start[q][c];
                uint16_t* cl = &L.cl_pos[size_t(b) * 2 * L.Hpad];
                for (uint32_t i = 0; i < L.n_live[b]; ++i) {  // ascending position = ascending strength on the river
                    for (int w = 0; w < 2; ++w) {
                        const int c = P->hand_cards[q][2 * sop[i] + w];
                        cl[fill[c]++] = uint16_t(i);
                        if (final_round) seg_str[q][c].push_back(P->sd[q].strength[size_t(b) * H + sop[i]]);
                    }
                }
                // boundary flags and per-thread first ordinals (tasks.h: cl_pos entry format)
                uint32_t n_lists = 0;
                for (int c = 0; c < 52; ++c) {
                    seg_ord[q][c] = -1;
                    if (seg_start[q][c + 1] > seg_start[q][c]) {
                        seg_ord[q][c] = int(n_lists++);
                        cl[seg_start[q][c]] |= uint16_t(CL_FIRST);
                    }
                }
                {
                    uint32_t started = 0;  // lists that start before the current entry
                    const uint32_t n_entries = 2 * L.Hpad;
                    for (uint32_t e = 0; e < n_entries; e += 8) {
                        const uint32_t ord = e == 0 ? n_lists : started;
                        cl[e] |= uint16_t((ord & 7u) << 11);
                        cl[e + 1] |= uint16_t(((ord >> 3) & 7u) << 11);
                        for (uint32_t x = e; x < e + 8 && x < n_entries; ++x)
                            if (cl[x] & CL_FIRST) ++started;
                    }
                }
                if (k > 0) {
                    const LocalTables& Lp = P->loc[k - 1][q];
                    const uint32_t pb = uint32_t(P->board_parent[k][b]);
                    uint16_t* pp = &L.parent_pos[size_t(b) * L.Hpad];
                    uint16_t* cp = &L.child_pos[size_t(b) * Lp.Hpad];
                    for (uint32_t i = 0; i < L.n_live[b]; ++i) {
                        const uint16_t ppos = Lp.pos_of_slot[size_t(pb) * H + sop[i]];
                        pp[i] = ppos;
                        cp[ppos] = uint16_t(i);
                    }
                }
            }
            for (int q = 0; q < 2; ++q) {
                const int o = 1 - q;
                LocalTables& L = P->loc[k][q];
                const LocalTables& Lo = P->loc[k][o];
                const uint32_t H = P->H[q], Ho = P->H[o];
                const uint16_t* sop = &L.slot_of_pos[size_t(b) * L.Hpad];
                std::vector<uint32_t> ostr;  // opponent strengths in its local (sorted) order
                if (final_round) {
                    ostr.resize(Lo.n_live[b]);
                    for (uint32_t i = 0; i < Lo.n_live[b]; ++i)
                        ostr[i] = P->sd[o].strength[size_t(b) * Ho + Lo.slot_of_pos[size_t(b) * Lo.Hpad + i]];
                }
                for (uint32_t i = 0; i < L.n_live[b]; ++i) {
                    const uint32_t h = sop[i];
                    const int c0 = P->hand_cards[q][2 * h], c1 = P->hand_cards[q][2 * h + 1];
                    uint32_t lo = 0, hi = 0, ab[2][2] = {{0, 0}, {0, 0}}, kk[2] = {HREC_K_NONE, HREC_K_NONE}, same4 = HREC_SAME_NONE;
                    const int cc[2] = {c0, c1};
                    const uint16_t sm = P->same[q][h];
                    if (sm != 0xFFFF) {
                        const uint16_t op = Lo.pos_of_slot[size_t(b) * Ho + sm];
                        if (op < Lo.n_live[b]) same4 = uint32_t(op) * 4u;
                    }
                    const uint32_t st = final_round ? P->sd[q].strength[size_t(b) * H + h] : 0;
                    if (final_round) {
                        lo = uint32_t(std::lower_bound(ostr.begin(), ostr.end(), st) - ostr.begin());
                        hi = uint32_t(std::upper_bound(ostr.begin(), ostr.end(), st) - ostr.begin());
                    }
                    for (int w = 0; w < 2; ++w) {
                        const int c = cc[w];
                        if (seg_ord[o][c] < 0) continue;  // the opponent holds no hand with this card: GB[0] = 0 and B[54..55] = 0
                        kk[w] = uint32_t(seg_ord[o][c]) * 4u;
                        const uint32_t s0 = seg_start[o][c];
                        uint32_t dlo = 0, dhi = 0;
                        if (final_round) {
                            const auto& l = seg_str[o][c];
                            dlo = uint32_t(std::lower_bound(l.begin(), l.end(), st) - l.begin());
                            dhi = uint32_t(std::upper_bound(l.begin(), l.end(), st) - l.begin());
                        }
                        ab[w][0] = (s0 + dlo) * 4u;
                        ab[w][1] = (s0 + dhi) * 4u;
                    }
                    HandRec r;
                    r.w0 = (lo * 4u) | ((hi * 4u) << 16);
                    r.w1 = ab[0][0] | (ab[0][1] << 16);
                    r.w2 = ab[1][0] | (ab[1][1] << 16);
                    r.w3 = kk[0] | (kk[1] << 9) | (same4 << 18);
                    L.hrec[size_t(b) * L.Hpad + i] = r;
                }
            }
        }
    }

    // final round: the hand's cards by board-local position, and whether both players' hands share positions
    if (any_showdown) {
        const uint32_t k = P->n_rounds - 1;
        const uint32_t nB = P->n_boards[k];
        bool same = P->H[0] == P->H[1];
        for (int q = 0; q < 2; ++q) {
            LocalTables& L = P->loc[k][q];
            L.pcards.assign(size_t(nB) * L.Hpad, 0);
            for (uint32_t b = P->local_lo[k]; b < P->local_hi[k]; ++b) {
                const uint16_t* sop = &L.slot_of_pos[size_t(b) * L.Hpad];
                for (uint32_t i = 0; i < L.n_live[b]; ++i) {
                    L.pcards[size_t(b) * L.Hpad + i] = uint16_t(P->hand_cards[q][2 * sop[i]] | (P->hand_cards[q][2 * sop[i] + 1] << 8));
                    if ((L.hrec[size_t(b) * L.Hpad + i].w3 >> 18) != i * 4u) same = false;  // same4: the identical combo sits at the same position
                }
                if (L.n_live[b] != P->loc[k][1 - q].n_live[b]) same = false;
            }
        }
        P->same_order = same;
    }

    // ---- task graphs, one per traverser ----
    for (int p = 0; p < 2; ++p) {
        TaskGen g{P, p};
        if (!g.run()) return fail(g.err);
        P->tl[p] = std::move(g.tl);
        P->street[p] = std::move(g.sp);
    }

    // ---- update counts (one update = one (node, board, row, action) cell) ----
    for (uint32_t k = 0; k < P->n_rounds; ++k)
        for (int q = 0; q < 2; ++q) {
            const RoundPlayerTables& T = P->tabs[k][q];
            for (uint32_t b = P->local_lo[k]; b < P->local_hi[k]; ++b)
                P->updates_per_iter_local += uint64_t(T.n_rows[b]) * T.sum_a;
        }
    P->updates_per_iter_global = P->updates_per_iter_local;  // engine all-reduces this when sharded
    return true;
}

void materialize_tasks(const Plan& P, int trav, const uint32_t counts[3], MaterializedTasks* out) {
    const TaskList& tl = P.tl[trav];
    std::vector<NodeTask>& t = out->tasks;
    std::vector<TaskSrc>& sr = out->srcs;
    t = tl.tasks;
    sr = tl.srcs;
    uint32_t first = 0;
    out->phase_cut = 0;
    bool cut_set = false;
    for (size_t j = 0; j < t.size(); ++j) {
        if (!cut_set && tl.tasks[j].first >= tl.phase_cut) {
            out->phase_cut = first;
            cut_set = true;
        }
        t[j].first = first;
        t[j].count = counts[t[j].round_k];
        first += t[j].count;
    }
    if (!cut_set) out->phase_cut = first;
    out->n_tickets = first;
    for (NodeTask& x : t)
        for (int i = 0; i < x.n_dep; ++i)
            if (x.dep[i] >= 0) x.dep[i] = int32_t(t[x.dep[i]].first);
    for (TaskSrc& x : sr)
        if (x.dep >= 0) x.dep = int32_t(t[x.dep].first);
}

bool build_execution_order(const Plan& P, int trav, const MaterializedTasks& m, const uint32_t counts[3], bool force, std::vector<uint32_t>* ord,
                           std::string* err) {
    ord->clear();
    const std::vector<NodeTask>& t = m.tasks;
    const uint32_t kf = P.n_rounds - 1;
    if (P.n_rounds < 2 || counts[kf] != P.boards_local(kf)) return true;
    const TaskList& tlp = P.tl[trav];
    const uint64_t vec_bytes = (uint64_t(tlp.n_rbuf[kf]) + tlp.n_cbuf[kf]) * counts[kf] * std::max(P.H[0], P.H[1]) * 4;
    uint32_t lo = UINT32_MAX, hi = 0;
    std::vector<uint32_t> fin;  // node tasks of the final round, list order (downs by depth, then ups deepest first)
    for (size_t j = 0; j < t.size(); ++j)
        if (t[j].round_k == kf && t[j].count) {
            fin.push_back(uint32_t(j));
            lo = std::min(lo, t[j].first);
            hi = std::max(hi, t[j].first + t[j].count);
        }
    uint64_t covered = 0;
    for (uint32_t j : fin) covered += t[j].count;
    const bool contiguous = !fin.empty() && covered == uint64_t(hi - lo);
    if (!contiguous || !(vec_bytes > (96ull << 20) || force)) return true;
    ord->resize(m.n_tickets);
    for (uint32_t i = 0; i < m.n_tickets; ++i) (*ord)[i] = i;
    const int32_t* par = P.board_parent[kf].data() + P.local_lo[kf];
    uint32_t pos = lo;
    for (uint32_t b0 = 0; b0 < counts[kf];) {
        uint32_t b1 = b0 + 1;
        while (b1 < counts[kf] && par[b1] == par[b0]) ++b1;
        for (uint32_t j : fin)
            for (uint32_t b = b0; b < b1; ++b) (*ord)[pos++] = t[j].first + b;
        b0 = b1;
    }
    if (pos != hi) {
        if (err) *err = "internal: execution order does not cover the final round";
        return false;
    }
    return true;
}

std::string check_execution_order(const Plan& P, int trav, const MaterializedTasks& m, const uint32_t counts[3], const std::vector<uint32_t>& ord) {
    (void)trav;
    const uint32_t n = m.n_tickets;
    std::vector<uint32_t> pos(n);
    if (ord.empty()) {
        for (uint32_t i = 0; i < n; ++i) pos[i] = i;
    } else {
        if (ord.size() != n) return "order has the wrong length";
        std::vector<uint8_t> seen(n, 0);
        for (uint32_t tk = 0; tk < n; ++tk) {
            if (ord[tk] >= n || seen[ord[tk]]) return "order is not a permutation (ticket " + std::to_string(tk) + ")";
            seen[ord[tk]] = 1;
            pos[ord[tk]] = tk;
        }
    }
    auto bad = [&](uint32_t producer, uint32_t slot, size_t j) -> std::string {
        if (producer >= n) return "task " + std::to_string(j) + ": producer slot " + std::to_string(producer) + " out of range";
        if (pos[producer] >= pos[slot])
            return "task " + std::to_string(j) + " slot " + std::to_string(slot) + " (ticket " + std::to_string(pos[slot]) + ") runs before its producer slot " +
                   std::to_string(producer) + " (ticket " + std::to_string(pos[producer]) + ")";
        return "";
    };
    for (size_t j = 0; j < m.tasks.size(); ++j) {
        const NodeTask& st = m.tasks[j];
        const uint32_t k = st.round_k;
        for (uint32_t inst = 0; inst < st.count; ++inst) {
            const uint32_t slot = st.first + inst, b = inst;
            if (st.kind == TK_GATHER) {
                const bool sharded_next = P.world > 1 && k + 1 == P.shard_round;
                const uint32_t per_parent = sharded_next ? 0u : P.deal_count[k + 1];
                const uint32_t gfirst = uint32_t(st.dep[0]) + (per_parent > 0 ? b * per_parent : 0u);
                const uint32_t total = per_parent > 0 ? per_parent : counts[k + 1];
                for (uint32_t c = 0; c < total; ++c) {
                    const std::string e = bad(gfirst + c, slot, j);
                    if (!e.empty()) return e;
                }
                continue;
            }
            for (int i = 0; i < st.n_dep; ++i) {
                if (st.dep[i] < 0) continue;
                uint32_t off = inst;
                if (st.dep_kind[i] == DK_PARENT_BOARD) off = uint32_t(P.board_parent[k][P.local_lo[k] + b] - int32_t(P.local_lo[k - 1]));
                const std::string e = bad(uint32_t(st.dep[i]) + off, slot, j);
                if (!e.empty()) return e;
            }
            for (uint32_t s = 0; s < st.n_src_all; ++s) {
                const int32_t d = m.srcs[st.src_all_first + s].dep;
                if (d < 0) continue;
                const std::string e = bad(uint32_t(d) + inst, slot, j);
                if (!e.empty()) return e;
            }
        }
    }
    return "";
}

}  // namespace rs
