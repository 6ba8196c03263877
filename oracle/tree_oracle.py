"""Pure-Python restatement of RustSolver's betting rules and tree builder.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module; the
product (rustsolver_b200) has its own C++ tree builder and never routes through here.

Follows, line for line:
    src/solver/constants.rs:1-6         ALLIN_THRESHOLD, MAX_RAISES, MAX_PLAYERS
    src/solver/state.rs:52-212          GameState::{from, is_terminal, is_allin, to_next_street,
                                        valid_actions, apply_action}
    src/solver/tree_builder.rs:60-143   build_private_chance / build_action_nodes / build_action /
                                        build_terminal / build_public_chance (DFS pre-order arena)
    src/solver/cfr.rs:73-98             generate_all_hole_card_combos (count only)

Pinned against SURVEY.md Appendix A (golden tree of options::default_flop()) and Appendix B (node
counts of the BASELINE.json configs) in tests/test_tree.py.  Those goldens were derived by the
survey from the same Rust sources (the reference ships no tree fixture), so: parity unpinned
against a reference RUN, pinned against two independent restatements.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, replace
from typing import List, Tuple

ALLIN_THRESHOLD = 0.67  # constants.rs:2
MAX_RAISES = 2          # constants.rs:5
MAX_PLAYERS = 2         # constants.rs:6

FLOP, TURN, RIVER = 0, 1, 2
BET, RAISE, CHECK, CALL, FOLD = 0, 1, 2, 3, 4
NODE_ACTION, NODE_TERMINAL, NODE_PUBLIC_CHANCE, NODE_PRIVATE_CHANCE = 0, 1, 2, 3
TERM_ALLIN, TERM_SHOWDOWN, TERM_UNCONTESTED = 0, 1, 2

U32_MAX = 0xFFFFFFFF


def f64_as_u32(x: float) -> int:
    """Rust `as u32` on an f64: truncate toward zero, saturate, NaN -> 0."""
    if math.isnan(x) or x <= 0.0:
        return 0
    if x >= float(U32_MAX):
        return U32_MAX
    return int(x)


@dataclass
class PlayerState:  # state.rs:24-41
    stack: int
    wager: int = 0
    has_folded: bool = False


@dataclass
class GameState:  # state.rs:43-50
    players: List[PlayerState]
    pot: int
    raise_count: int
    current: int
    round: int
    bets_settled: bool

    def clone(self) -> "GameState":
        return GameState([replace(p) for p in self.players], self.pot, self.raise_count, self.current,
                         self.round, self.bets_settled)

    @staticmethod
    def from_options(stack_sizes, board_mask: int, starting_pot: int) -> "GameState":  # state.rs:52-71
        n = bin(board_mask).count("1")
        if n not in (3, 4, 5):
            raise ValueError("invalid board mask")
        return GameState([PlayerState(stack_sizes[0]), PlayerState(stack_sizes[1])], starting_pot, 0, 0,
                         {3: FLOP, 4: TURN, 5: RIVER}[n], False)

    def is_uncontested(self) -> bool:  # state.rs:86-93
        return any(p.has_folded for p in self.players)

    def is_allin(self) -> bool:  # state.rs:99-106
        return any(p.stack == 0 for p in self.players)

    def is_terminal(self) -> bool:  # state.rs:94-98
        return self.round == RIVER or self.is_allin() or self.is_uncontested()

    def to_next_street(self) -> "GameState":  # state.rs:107-123 (raise_count is not reset)
        n = self.clone()
        n.bets_settled = False
        n.current = 0
        for p in n.players:
            p.wager = 0
        if self.round == FLOP:
            n.round = TURN
        elif self.round == TURN:
            n.round = RIVER
        else:
            raise RuntimeError("Should not get here")
        return n

    def valid_actions(self, bet_sizes, raise_sizes, round_idx: int) -> List[Tuple[int, float]]:  # state.rs:124-156
        me, other = self.players[self.current], self.players[1 - self.current]
        actions: List[Tuple[int, float]] = []
        if other.wager == 0:
            actions.append((CHECK, 0.0))
        if other.wager > me.wager:
            actions.append((CALL, 0.0))
        if other.wager > me.wager:
            actions.append((FOLD, 0.0))
        if other.wager == 0:
            for bet_size in bet_sizes[round_idx]:
                chips = bet_size * float(self.pot)
                actions.append((BET, bet_size))
                if chips > ALLIN_THRESHOLD * float(me.stack):
                    break
        if self.raise_count < MAX_RAISES and not self.is_allin() and other.wager > me.wager:
            for raise_size in raise_sizes[round_idx]:
                chips = raise_size * float(other.wager)
                actions.append((RAISE, raise_size))
                if chips > ALLIN_THRESHOLD * float(me.stack):
                    break
        return actions

    def apply_action(self, action: Tuple[int, float]) -> "GameState":  # state.rs:157-212
        kind, amt = action
        n = self.clone()
        me = n.players[n.current]
        other_wager = self.players[1 - self.current].wager
        if kind == BET:
            chips = f64_as_u32(float(n.pot) * amt)
            if chips > f64_as_u32(float(me.stack) * ALLIN_THRESHOLD):
                chips = me.stack
            me.stack -= chips
            me.wager = chips
            n.pot += chips
            n.current = 1 - n.current
        elif kind == RAISE:
            chips = f64_as_u32(float(other_wager) * amt)
            if chips > f64_as_u32(float(me.stack) * ALLIN_THRESHOLD):
                chips = me.stack
            me.stack -= chips
            me.wager += chips
            n.raise_count += 1
            n.pot += chips
            n.current = 1 - n.current
        elif kind == CALL:
            wager_diff = other_wager - me.wager
            if me.stack >= wager_diff:
                n.pot += wager_diff
                me.stack -= wager_diff
            else:
                n.pot += me.stack
                me.stack = 0
            n.bets_settled = True  # the player does not switch
        elif kind == CHECK:
            if n.current == MAX_PLAYERS - 1:
                n.bets_settled = True
            n.current = 1 - n.current
        elif kind == FOLD:
            me.has_folded = True
            wager_diff = other_wager - me.wager
            n.pot -= wager_diff
            n.bets_settled = True
        return n


@dataclass
class Node:
    type: int
    parent: int
    children: List[int] = field(default_factory=list)
    actions: List[Tuple[int, float]] = field(default_factory=list)
    index: int = 0
    player: int = 0
    round_idx: int = 0
    value: int = 0
    ttype: int = TERM_SHOWDOWN
    last_to_act: int = 0
    round: int = RIVER


class TreeBuilder:  # tree_builder.rs:16-143
    def __init__(self, bet_sizes, raise_sizes):
        self.nodes: List[Node] = []
        self.n_actions = 0
        self.bet_sizes, self.raise_sizes = bet_sizes, raise_sizes

    def create_node(self, node: Node) -> int:  # tree.rs:48-53
        self.nodes.append(node)
        return len(self.nodes) - 1

    def build_private_chance(self, state: GameState):  # :60-66
        node = self.create_node(Node(NODE_PRIVATE_CHANCE, -1))
        child = self.build_action_nodes(node, 0, state)
        self.nodes[node].children.append(child)

    def build_action_nodes(self, parent: int, round_idx: int, state: GameState) -> int:  # :67-90
        node_id = self.create_node(Node(NODE_ACTION, parent, player=state.current, index=self.n_actions,
                                        round_idx=round_idx))
        self.n_actions += 1
        for action in state.valid_actions(self.bet_sizes, self.raise_sizes, round_idx):
            self.build_action(node_id, round_idx, state, action)
        return node_id

    def build_action(self, node: int, round_idx: int, state: GameState, action):  # :91-115
        next_state = state.apply_action(action)
        if next_state.bets_settled:
            if next_state.is_terminal():
                child = self.build_terminal(node, next_state)
            else:
                child = self.build_public_chance(node, round_idx, next_state.to_next_street())
        else:
            child = self.build_action_nodes(node, round_idx, next_state)
        self.nodes[node].children.append(child)
        self.nodes[node].actions.append(action)

    def build_terminal(self, parent: int, state: GameState) -> int:  # :116-133
        t = Node(NODE_TERMINAL, parent, value=state.pot, ttype=TERM_SHOWDOWN, last_to_act=state.current,
                 round=state.round)
        if state.is_allin() and state.round != RIVER:
            t.ttype = TERM_ALLIN
        if state.is_uncontested():
            t.ttype = TERM_UNCONTESTED
        return self.create_node(t)

    def build_public_chance(self, parent: int, round_idx: int, state: GameState) -> int:  # :134-143
        node = self.create_node(Node(NODE_PUBLIC_CHANCE, parent, round=state.round))
        child = self.build_action_nodes(node, round_idx + 1, state)
        self.nodes[node].children.append(child)
        return node


def build_game_tree(stack_sizes, board_mask: int, starting_pot: int, bet_sizes, raise_sizes):
    """tree_builder.rs:9-14 -> (n_actions, [Node])."""
    b = TreeBuilder(bet_sizes, raise_sizes)
    b.build_private_chance(GameState.from_options(stack_sizes, board_mask, starting_pot))
    return b.n_actions, b.nodes


def dump(nodes: List[Node]) -> List[str]:
    """SURVEY.md Appendix A notation."""
    out = []
    for i, n in enumerate(nodes):
        if n.type == NODE_PRIVATE_CHANCE:
            out.append(f"{i} P ->{n.children}")
        elif n.type == NODE_PUBLIC_CHANCE:
            out.append(f"{i} C ->{n.children}")
        elif n.type == NODE_TERMINAL:
            k = {TERM_ALLIN: "L", TERM_SHOWDOWN: "S", TERM_UNCONTESTED: "U"}[n.ttype]
            out.append(f"{i} {k} {n.value}/{n.last_to_act}")
        else:
            acts = [{CHECK: "X", CALL: "C", FOLD: "F"}.get(k) or (("B" if k == BET else "R") + f"{a:g}") for k, a in n.actions]
            out.append(f"{i} A {n.index}/P{n.player} [{','.join(acts)}] ->{n.children}")
    return out


def flatten(nodes: List[Node]):
    """-> dict of plain lists in the rs_tree layout (include/b200cfr.h)."""
    child_offset, children = [0], []
    for n in nodes:
        children.extend(n.children)
        child_offset.append(len(children))
    return dict(
        type=[n.type for n in nodes], parent=[n.parent for n in nodes], child_offset=child_offset,
        children=children, player=[n.player for n in nodes], an_index=[n.index for n in nodes],
        round_idx=[n.round_idx for n in nodes], value=[n.value for n in nodes], ttype=[n.ttype for n in nodes],
        last_to_act=[n.last_to_act for n in nodes], round=[n.round for n in nodes],
        action_kind=[k for n in nodes for (k, _) in (n.actions if n.type == NODE_ACTION else [(0xFF, 0.0)] * len(n.children))],
        action_amount=[a for n in nodes for (_, a) in (n.actions if n.type == NODE_ACTION else [(0xFF, 0.0)] * len(n.children))],
    )


def tree_stats(nodes: List[Node]) -> dict:
    """Counts in the layout of SURVEY.md Appendix B."""
    s = dict(nodes=len(nodes), action_nodes=0, infoset_actions=0, chance=0, showdown=0, fold=0, allin=0,
             action_nodes_per_round={}, infoset_actions_per_round={})
    for n in nodes:
        if n.type == NODE_ACTION:
            s["action_nodes"] += 1
            s["infoset_actions"] += len(n.children)
            s["action_nodes_per_round"][n.round_idx] = s["action_nodes_per_round"].get(n.round_idx, 0) + 1
            s["infoset_actions_per_round"][n.round_idx] = s["infoset_actions_per_round"].get(n.round_idx, 0) + len(n.children)
        elif n.type == NODE_PUBLIC_CHANCE:
            s["chance"] += 1
        elif n.type == NODE_TERMINAL:
            s[{TERM_ALLIN: "allin", TERM_SHOWDOWN: "showdown", TERM_UNCONTESTED: "fold"}[n.ttype]] += 1
    return s
