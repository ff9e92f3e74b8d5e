class _NoOp(object):
    """Callable, attribute-chaining, context-manager singleton (fjcommon.no_op.NoOp)."""
    def __getattr__(self, _):
        return self

    def __call__(self, *a, **kw):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


NoOp = _NoOp()
