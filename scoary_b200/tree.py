"""Nested-list trees, the form Scoary's hot path consumes.

The reference represents the (UPGMA or user) tree as nested 2-lists of isolate
names, e.g. [["a", "b"], "c"] (scoary/methods.py:667-707 builds it,
:709-739 prunes it, :741-752 writes it, :1386-1402 walks it).  The engine takes
the same tree flattened into child-index arrays (include/scoary_b200.h,
sb_set_tree).  Everything here is iterative: UPGMA trees on real data are deep
(height ~N/8) and recursion would hit Python's limit long before N = 10 000.
"""
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


def _resolve_polytomy(children):
    """ete3's TreeNode.resolve_polytomy(recursive=True), which the reference applies to every custom tree
    (scoary/nwkhandler.py:19): a node with children c0 .. c(k-1), k > 2, becomes the ladder
    [[...[[c(k-2), c(k-1)], c(k-3)] ..., c1], c0].  A node with one child is that child."""
    if len(children) == 1:
        return children[0]
    node = [children[-2], children[-1]]
    for ch in reversed(children[:-2]):
        node = [node, ch]
    return node


def from_newick(text):
    """Newick text -> nested 2-lists of leaf names, as nwkhandler.ReadTreeFromFile / RecTree2List build them
    (scoary/nwkhandler.py:10-40): branch lengths, internal node labels / support values and [comments] are
    discarded, quotes around leaf names are stripped, polytomies (e.g. the trifurcating root of an unrooted ML
    tree) are resolved like ete3 does.  Also reads the files Scoary writes itself (to_scoary_newick: quoted names,
    no lengths).  Iterative: UPGMA trees are deep.  Raises ValueError on malformed input."""
    body = text.strip()
    n = len(body)
    seps = ",();"

    def skip_meta(i):       # label, ':length', [comment] after a name or a ')': up to the next structural character
        while i < n and body[i] not in seps:
            if body[i] == "[":
                j = body.find("]", i + 1)
                if j < 0:
                    raise ValueError("unterminated [comment]")
                i = j + 1
            elif body[i] in "'\"":      # a quoted internal label
                j = body.find(body[i], i + 1)
                if j < 0:
                    raise ValueError("unterminated quoted name")
                i = j + 1
            else:
                i += 1
        return i

    stack, root, i = [], None, 0
    while i < n:
        ch = body[i]
        if ch == "(":
            stack.append([])
            i += 1
        elif ch == ")":
            if not stack:
                raise ValueError("unbalanced parentheses")
            kids = stack.pop()
            if not kids:
                raise ValueError("empty node")
            node = _resolve_polytomy(kids)
            i = skip_meta(i + 1)
            if stack:
                stack[-1].append(node)
            else:
                root = node
                break
        elif ch in ", \t\r\n":
            i += 1
        elif ch == ";":
            break
        elif ch == "[":
            i = skip_meta(i)
        else:
            if ch in "'\"":
                j = body.find(ch, i + 1)
                if j < 0:
                    raise ValueError("unterminated quoted name")
                name, i = body[i + 1:j], j + 1
            else:           # bare name up to : , ( ) ; [
                j = i
                while j < n and body[j] not in ",():;[":
                    j += 1
                name, i = body[i:j].strip(), j
            i = skip_meta(i)
            if not stack:
                raise ValueError("a tree needs at least two leaves")
            stack[-1].append(name.lstrip("'\"").rstrip("'\""))     # nwkhandler.py:33
    if root is None or stack:
        raise ValueError("unbalanced parentheses")
    if body[i:].strip().lstrip(";").strip():
        raise ValueError("text after the end of the tree")
    if isinstance(root, str):
        raise ValueError("a tree needs at least two leaves")
    return root


from_scoary_newick = from_newick


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
