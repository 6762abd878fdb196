
import sys, types, importlib.machinery
class _Any:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Any()
    def __getattr__(self, n):
        if n.startswith('__') and n.endswith('__'): raise AttributeError(n)
        return _Any()
    def __iter__(self): return iter(())
    def __mro_entries__(self, bases): return (object,)
    def __class_getitem__(cls, item): return cls
    def __or__(self, other): return self
    def __ror__(self, other): return self
class _Mod(types.ModuleType):
    __path__ = []
    def __getattr__(self, n):
        if n.startswith('__') and n.endswith('__'): raise AttributeError(n)
        full = self.__name__ + '.' + n
        if full in sys.modules: return sys.modules[full]
        if n[:1].isupper() or n in ('is_compiled_module',):
            return type(n, (object,), {'__init__': lambda s,*a,**k: None, '__getattr__': lambda s, k: _Any()})
        return _Any()
class _FFB200StubFinder:
    ROOTS = ('accelerate', 'peft', 'imageio')
    def find_spec(self, name, path=None, target=None):
        root = name.split('.')[0]
        if root in self.ROOTS and name != root:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None
    def create_module(self, spec):
        m = _Mod(spec.name); m.__path__ = []; return m
    def exec_module(self, module): pass
if not any(type(f).__name__ == '_FFB200StubFinder' for f in sys.meta_path):
    sys.meta_path.append(_FFB200StubFinder())
_self = sys.modules[__name__]
_self.__class__ = _Mod
__version__ = '0.17.0'
