"""Message / reduce descriptors (``dgl.function`` stand-in; test infrastructure only)."""


class _Fn:
    def __init__(self, kind, lhs=None, rhs=None, out=None):
        self.kind, self.lhs, self.rhs, self.out = kind, lhs, rhs, out


def u_mul_e(lhs, rhs, out):
    return _Fn('u_mul_e', lhs, rhs, out)


def u_add_v(lhs, rhs, out):
    return _Fn('u_add_v', lhs, rhs, out)


def copy_u(u, out):
    return _Fn('copy_u', u, None, out)


def copy_src(src, out):
    return _Fn('copy_u', src, None, out)


def sum(msg, out):  # noqa: A001 - mirrors dgl.function.sum
    return _Fn('sum', msg, None, out)
