"""Re-export a pointcloudlib_b200.misc module with outputs converted to jittor-compat Vars."""
import functools
import inspect

import torch


def _to_var(x):
    from jittor import _wrap
    return _wrap(x)


def export(src_module, namespace):
    for name, obj in vars(src_module).items():
        if name.startswith("_"):
            continue
        if inspect.isclass(obj) and issubclass(obj, torch.nn.Module) and obj.__module__ == src_module.__name__:
            def _call(self, *a, __base=obj, **k):
                return _to_var(__base.__call__(self, *a, **k))
            namespace[name] = type(name, (obj,), {"__call__": _call, "__module__": namespace["__name__"],
                                                  "__doc__": obj.__doc__})
        elif inspect.isfunction(obj) and obj.__module__ == src_module.__name__:
            def _fn(*a, __f=obj, **k):
                return _to_var(__f(*a, **k))
            namespace[name] = functools.wraps(obj)(_fn)
        else:
            namespace[name] = obj
