import types
checkify = types.SimpleNamespace(checkify=lambda f, *a, **k: f, check=lambda *a, **k: None)


def io_callback(cb, result_shape, *args, **kw):
    return cb(*args)
