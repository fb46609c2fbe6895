"""jax.tree_util stand-in: dict / list / tuple / None / registered nodes.  Test infrastructure only."""

_REG = {}


def register_pytree_node(cls, flatten, unflatten):
    _REG[cls] = (flatten, unflatten)


def tree_map(f, tree, *rest, is_leaf=None):
    if is_leaf is not None and is_leaf(tree):
        return f(tree, *rest)
    t = type(tree)
    if t in _REG:
        fl, un = _REG[t]
        ch, aux = fl(tree)
        rch = [_REG[t][0](r)[0] for r in rest]
        return un(aux, [tree_map(f, c, *[r[i] for r in rch], is_leaf=is_leaf) for i, c in enumerate(ch)])
    if isinstance(tree, dict):
        return {k: tree_map(f, v, *[r[k] for r in rest], is_leaf=is_leaf) for k, v in tree.items()}
    if isinstance(tree, tuple) and hasattr(tree, "_fields"):
        return t(*[tree_map(f, v, *[r[i] for r in rest], is_leaf=is_leaf) for i, v in enumerate(tree)])
    if isinstance(tree, (list, tuple)):
        return t(tree_map(f, v, *[r[i] for r in rest], is_leaf=is_leaf) for i, v in enumerate(tree))
    if tree is None:
        return None
    return f(tree, *rest)


def tree_leaves(tree, is_leaf=None):
    out = []
    tree_map(lambda x: out.append(x), tree, is_leaf=is_leaf)
    return out


def tree_flatten(tree, is_leaf=None):
    leaves = tree_leaves(tree, is_leaf)
    return leaves, tree


def tree_unflatten(treedef, leaves):
    it = iter(leaves)
    return tree_map(lambda _: next(it), treedef)


map = tree_map
leaves = tree_leaves
flatten = tree_flatten
unflatten = tree_unflatten
