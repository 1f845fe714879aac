"""Minimal stand-in for chainer 4.x (tests only; see ../README.md): Variable + the FunctionNode protocol."""
import heapq

from chainer import configuration, cuda, function_node, utils            # noqa: F401
from chainer.configuration import config, using_config                   # noqa: F401
from chainer.function_node import FunctionNode                           # noqa: F401
from chainer.variable import Variable                                    # noqa: F401
from chainer import functions                                            # noqa: F401,E402

__standin__ = True
__version__ = "4.1.0-standin"


def backward_all(seeds):
    """Backpropagate from several Variables whose ``.grad`` is set (what ``loss.backward()`` does from one scalar)."""
    heap, seen, order = [], set(), 0
    for v in seeds:
        n = v.creator_node
        if n is not None and id(n) not in seen:
            seen.add(id(n))
            heapq.heappush(heap, (-n.rank, order, n))
            order += 1
    while heap:
        _, _, node = heapq.heappop(heap)
        outs = [o() for o in node.outputs]
        gys = tuple(None if (o is None or o.grad is None) else Variable(o.grad, requires_grad=False) for o in outs)
        idx = tuple(i for i, v in enumerate(node.inputs) if v.requires_grad)
        if not idx or all(g is None for g in gys):
            continue
        gxs = node.backward(idx, gys)
        if len(gxs) == len(node.inputs):
            gxs = tuple(gxs[i] for i in idx)
        assert len(gxs) == len(idx), "backward returned %d gradients for %d requested inputs" % (len(gxs), len(idx))
        for i, g in zip(idx, gxs):
            if g is None:
                continue
            v = node.inputs[i]
            g = g.data if isinstance(g, Variable) else g
            v.grad = g if v.grad is None else v.grad + g
            n = v.creator_node
            if n is not None and id(n) not in seen:
                seen.add(id(n))
                heapq.heappush(heap, (-n.rank, order, n))
                order += 1
