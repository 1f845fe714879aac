class Variable(object):
    def __init__(self, data=None, requires_grad=True, name=None):
        self.data = data
        self.grad = None                 # an array, as in chainer
        self.requires_grad = requires_grad
        self.creator_node = None
        self.rank = 0
        self.name = name

    array = property(lambda self: self.data)
    shape = property(lambda self: self.data.shape)
    dtype = property(lambda self: self.data.dtype)
    ndim = property(lambda self: self.data.ndim)

    def backward(self):
        import chainer
        assert self.grad is not None, "stand-in: set .grad before backward()"
        chainer.backward_all([self])

    def cleargrad(self):
        self.grad = None
