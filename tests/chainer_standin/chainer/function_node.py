import weakref

from chainer.utils import type_check
from chainer.variable import Variable


class FunctionNode(object):
    """chainer 4.x new-style function protocol: apply() -> check_type_forward / forward; backward(indexes, grad_outputs)."""
    inputs = None
    outputs = None
    rank = 0

    def check_type_forward(self, in_types):
        pass

    def forward(self, inputs):
        raise NotImplementedError

    def backward(self, target_input_indexes, grad_outputs):
        raise NotImplementedError

    def retain_inputs(self, indexes):
        self._input_indexes_to_retain = tuple(indexes)

    def get_retained_inputs(self):
        return tuple(self.inputs[i] for i in self._input_indexes_to_retain)

    def apply(self, inputs):
        in_vars = [x if isinstance(x, Variable) else Variable(x, requires_grad=False) for x in inputs]
        in_data = tuple(v.data for v in in_vars)
        self.check_type_forward(type_check.get_types(in_data, "in_types", False))
        self._input_indexes_to_retain = ()
        self.inputs = in_vars
        outs = self.forward(in_data)
        assert isinstance(outs, tuple), "forward must return a tuple"
        self.rank = max([v.rank for v in in_vars] + [0]) + 1
        need = any(v.requires_grad for v in in_vars)
        out_vars = []
        for o in outs:
            v = Variable(o, requires_grad=need)
            if need:
                v.creator_node = self
                v.rank = self.rank
            out_vars.append(v)
        self.outputs = [weakref.ref(v) for v in out_vars]
        self._keep = out_vars                 # the stand-in keeps the graph alive from both ends
        return tuple(out_vars)
