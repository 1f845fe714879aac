from chainer import function_node
from chainer.variable import Variable


class Reshape(function_node.FunctionNode):
    def __init__(self, shape):
        self.shape = shape

    def forward(self, inputs):
        x, = inputs
        self._in_shape = x.shape
        return x.reshape(self.shape),

    def backward(self, indexes, grad_outputs):
        return Variable(grad_outputs[0].data.reshape(self._in_shape), requires_grad=False),


def reshape(x, shape):
    return Reshape(shape).apply((x,))[0]
