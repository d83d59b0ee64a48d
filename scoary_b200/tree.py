"""Nested-list trees, the form Scoary's hot path consumes.

The reference represents the (UPGMA or user) tree as nested 2-lists of isolate
names, e.g. [["a", "b"], "c"] (scoary/methods.py:667-707 builds it,
:709-739 prunes it, :741-752 writes it, :1386-1402 walks it).  The engine takes
the same tree flattened into child-index arrays (include/scoary_b200.h,
sb_set_tree).  Everything here is iterative: UPGMA trees on real data are deep
(height ~N/8) and recursion would hit Python's limit long before N = 10 000.
"""
import ast

import numpy as np


def flatten(tree):
    """nested lists -> (left, right, leaf_names).

    Internal nodes are listed children-before-parents with the root last;
    a child >= 0 is an internal node index, a child < 0 is leaf id ~child; leaf
    ids follow the left-to-right order of the nested list."""
    if isinstance(tree, str) or tree is None or len(tree) != 2:
        raise ValueError("tree must be a nested list with at least two leaves")
    left, right, names = [], [], []
    stack = [[tree, 0, 0]]          # node, phase, ref of the left child
    ret = 0
    while stack:
        fr = stack[-1]
        node = fr[0]
        if isinstance(node, str):
            names.append(node)
            ret = ~(len(names) - 1)
            stack.pop()
        elif fr[1] == 0:
            if len(node) != 2:
                raise ValueError("tree is not binary")
            fr[1] = 1
            stack.append([node[0], 0, 0])
        elif fr[1] == 1:
            fr[2] = ret
            fr[1] = 2
            stack.append([node[1], 0, 0])
        else:
            left.append(fr[2])
            right.append(ret)
            ret = len(left) - 1
            stack.pop()
    return np.asarray(left, dtype=np.int32), np.asarray(right, dtype=np.int32), names


def leaves(tree):
    out, stack = [], [tree]
    while stack:
        n = stack.pop()
        if isinstance(n, str):
            out.append(n)
        elif n is not None:
            stack.append(n[1])
            stack.append(n[0])
    return out


def prune(tree, drop):
    """PruneForMissing (scoary/methods.py:709-739): remove the isolates in `drop`;
    a node left with one child is replaced by that child.  Returns None when
    nothing is left (the reference's bare `return`).

    The reference's UPGMA can leave a None where a cluster should be (when real distances tie
    with the 1 it assigns to dead clusters, methods.py:685-686); its Prunedic always ends with
    None (:612), so this pass is also what removes those: None children are dropped here."""
    drop = set(x for x in drop if x is not None)
    # post-order, iterative
    result = {}
    stack = [(tree, False)]
    while stack:
        node, done = stack.pop()
        if isinstance(node, str) or node is None:
            continue
        if not done:
            stack.append((node, True))
            stack.append((node[1], False))
            stack.append((node[0], False))
        else:
            kids = []
            for ch in node:
                if isinstance(ch, str):
                    kids.append(None if ch in drop else ch)
                elif ch is None:
                    kids.append(None)
                else:
                    kids.append(result.pop(id(ch)))
            if kids[0] is None and kids[1] is None:
                result[id(node)] = None
            elif kids[0] is None:
                result[id(node)] = kids[1]
            elif kids[1] is None:
                result[id(node)] = kids[0]
            else:
                result[id(node)] = [kids[0], kids[1]]
    if isinstance(tree, str) or tree is None:
        return None if (tree is None or tree in drop) else tree
    return result[id(tree)]


def to_scoary_newick(tree):
    """StoreUPGMAtreeToFile's text (scoary/methods.py:741-752): str(list) with
    brackets turned into parentheses, plus ';'."""
    parts, stack = [], [tree]
    # str() of nested lists, iteratively:  ['a', ['b', 'c']] -> "['a', ['b', 'c']]"
    while stack:
        n = stack.pop()
        if isinstance(n, tuple):       # literal text
            parts.append(n[0])
        elif isinstance(n, str) or n is None:
            parts.append(repr(n))
        else:
            stack.append(("]",))
            stack.append(n[1])
            stack.append((", ",))
            stack.append(n[0])
            stack.append(("[",))
    return "".join(parts).replace("[", "(").replace("]", ")") + ";"


def from_scoary_newick(text):
    """Inverse of to_scoary_newick for files Scoary wrote itself (quoted names,
    no branch lengths), e.g. exampledata/ExampleTree.nwk."""
    body = text.strip().rstrip(";").strip()
    # iterative parse of the restricted grammar: ( item , item ) with quoted leaves
    stack, cur, i, n = [], None, 0, len(body)
    root = None
    while i < n:
        ch = body[i]
        if ch == "(":
            new = []
            if cur is not None:
                cur.append(new)
                stack.append(cur)
            cur = new
            i += 1
        elif ch == ")":
            if len(cur) != 2:
                raise ValueError("tree is not binary (resolve polytomies first)")
            done = cur
            cur = stack.pop() if stack else None
            if cur is None:
                root = done
                break
            i += 1
        elif ch in "'\"":
            j = body.index(ch, i + 1)
            cur.append(ast.literal_eval(body[i:j + 1]))
            i = j + 1
        elif ch in ", \t\r\n":
            i += 1
        else:   # unquoted name up to , ) or :
            j = i
            while j < n and body[j] not in ",():":
                j += 1
            name = body[i:j].strip()
            if j < n and body[j] == ":":   # skip a branch length
                while j < n and body[j] not in ",)":
                    j += 1
            if name:            # an empty name is the branch length / label of a closed subtree
                cur.append(name)
            i = j
    if root is None:
        raise ValueError("could not parse tree")
    return root


def from_merges(names, merges):
    """UPGMA merge list [(i, j), ...] -> nested lists: the joined cluster keeps index i
    (scoary/methods.py:683,700-703)."""
    cluster = list(names)
    new = None
    for i, j in merges:
        i, j = int(i), int(j)
        new = [cluster[i], cluster[j]]
        cluster[i] = new
        cluster[j] = None
    return new


def random_join_tree(names, rng):
    """Seeded random-join (coalescent-shaped) binary tree used for the synthetic
    workloads with N >= 2000; rng is a numpy Generator."""
    nodes = list(names)
    while len(nodes) > 1:
        i = int(rng.integers(len(nodes)))
        a = nodes[i]
        nodes[i] = nodes[-1]
        nodes.pop()
        j = int(rng.integers(len(nodes)))
        b = nodes[j]
        nodes[j] = [a, b]
    return nodes[0]
