// Task graph of one CFR traversal, shared by the host plan compiler and the device kernel.
//
// A traversal of the public tree is cut into node-tasks; each node-task is
// instantiated once per board of its round, and one CTA executes one instance.  Instances are
// numbered in a topological order ("tickets"); a persistent kernel hands tickets out with one
// atomic counter and instances wait on their producers through per-instance flags.
//
//   TK_DOWN        opponent action node: sigma from the opponent's table, reach of every child
//                  (cfr.rs:582-586); terminal children are valued on the spot (cfr.rs:523-558) and
//                  summed into the node's terminal partial value
//   TK_UP_OPP      opponent action node: value = terminal partial + sum of child values
//   TK_UP_TRAV     traverser action node: child values, node value, regret and strategy-sum update
//                  (cfr.rs:588, 612-621)
//   TK_GATHER      chance node: sum of the child-street root values over the dealt cards
//                  (cfr.rs:502-522)
//   TK_ROOT_SHOWDOWN / TK_CHANCE_DOWN / TK_CHANCE_UP   street roots created by the all-in run-out
//                  expansion (a bare showdown, or a pass-through chance node)
#pragma once
#include <cstdint>

namespace rs {

enum TaskKind : uint8_t {
    TK_DOWN = 0,
    TK_UP_OPP = 1,
    TK_UP_TRAV = 2,
    TK_GATHER = 3,
    TK_ROOT_SHOWDOWN = 4,
    TK_CHANCE_DOWN = 5,
    TK_CHANCE_UP = 6
};
enum ChildKind : uint8_t { CK_ACTION = 0, CK_FOLD = 1, CK_SHOWDOWN = 2, CK_CHANCE = 3 };
enum DepKind : uint8_t { DK_NONE = 0, DK_SAME_BOARD = 1, DK_PARENT_BOARD = 2, DK_CHILD_BOARDS = 3 };

constexpr int32_t RIN_INITIAL = -1;  // reach source = the opponent's range weights (root round)
constexpr int MAX_TASK_CHILDREN = 8;
constexpr int MAX_TASK_DEPS = 10;
constexpr int MAX_TERMINAL_CHILDREN = 3;  // terminal children of one opponent node staged in shared memory

struct TaskChild {
    uint8_t kind;  // ChildKind
    uint8_t pad[3];
    int32_t buf;   // TK_DOWN: reach buffer written for this child; TK_UP_*: value buffer read (CK_ACTION) or leaf id (CK_CHANCE)
    float coef;    // +-pot of a terminal child (cfr.rs:525-556)
};

struct NodeTask {
    uint8_t kind;     // TaskKind
    uint8_t round_k;  // round_idx of the node
    uint8_t n_act;
    uint8_t n_dep;
    uint8_t rin_parent_round;  // 1: r_in names a reach buffer of the PARENT round, read at the parent board
    uint8_t pad0[3];
    int32_t r_in;     // reach buffer id, or RIN_INITIAL
    uint32_t cum_a;   // slab offset = n_rows(board) * cum_a inside the (round, player) table
    int32_t out;      // TK_DOWN: terminal-partial value buffer (-1: none); TK_UP_*, roots: value buffer; TK_GATHER: leaf id
    int32_t aux;      // TK_UP_OPP: terminal-partial buffer to add (-1 none); TK_GATHER: value buffer of the child street's root;
                      // TK_CHANCE_DOWN: reach buffer written; TK_CHANCE_UP: leaf id read
    uint32_t first;   // ticket of instance 0
    uint32_t count;   // instances = local boards of round_k
    uint32_t an_index;
    int32_t dep[MAX_TASK_DEPS];      // node-task index
    uint8_t dep_kind[MAX_TASK_DEPS]; // DepKind
    uint8_t pad1[2];
    TaskChild child[MAX_TASK_CHILDREN];
};
static_assert(sizeof(NodeTask) == 4 + 4 + 4 * 7 + 4 * MAX_TASK_DEPS + MAX_TASK_DEPS + 2 + 12 * MAX_TASK_CHILDREN, "NodeTask layout");

struct TaskCtl {  // device-resident dispatcher state, reset by the last CTA to leave
    unsigned long long ticket;
    unsigned int exited;
    unsigned int epoch;
};

}  // namespace rs
