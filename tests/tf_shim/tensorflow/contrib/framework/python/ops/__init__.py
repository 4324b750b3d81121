"""tf.contrib.framework arg_scope / add_arg_scope: keyword defaults injected into decorated functions."""
import contextlib
import functools

_stack = []


def add_arg_scope(fn):
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        merged = {}
        for scope in _stack:
            if wrapped in scope:
                merged.update(scope[wrapped])
        merged.update(kwargs)
        return fn(*args, **merged)
    wrapped._arg_scoped = True
    return wrapped


@contextlib.contextmanager
def arg_scope(fns, **kwargs):
    for f in fns:
        assert getattr(f, '_arg_scoped', False), '%r is not decorated with add_arg_scope' % f
    _stack.append({f: dict(kwargs) for f in fns})
    try:
        yield
    finally:
        _stack.pop()
