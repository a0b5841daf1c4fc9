"""Minimal stand-in for AllenNLP's Registrable / from_params so that the `model:` block of the
reference's expt/*/config.yaml instantiates unchanged (tell/commands/train.py:65-75).
Unknown keys raise, like Params.assert_empty."""
import inspect


class ConfigurationError(Exception):
    pass


class Registrable:
    _registry = {}

    @classmethod
    def register(cls, name):
        def deco(sub):
            Registrable._registry.setdefault(cls, {})[name] = sub
            return sub
        return deco

    @classmethod
    def registered_names(cls):
        """{name: class} registered on this base (or on bases related to it)."""
        out = {}
        for base, table in Registrable._registry.items():
            if issubclass(base, cls) or issubclass(cls, base):
                out.update(table)
        return out

    @classmethod
    def by_name(cls, name):
        for base, table in Registrable._registry.items():
            if issubclass(base, cls) or issubclass(cls, base):
                if name in table:
                    return table[name]
        raise ConfigurationError('%s is not a registered name for %s' % (name, cls.__name__))

    @classmethod
    def from_params(cls, params, **extras):
        params = dict(params)
        sub = cls.by_name(params.pop('type')) if 'type' in params else cls
        if hasattr(sub, '_from_params'):
            return sub._from_params(params, **extras)
        sig = inspect.signature(sub.__init__)
        kwargs = {}
        for name, p in list(sig.parameters.items())[1:]:
            if name in extras:
                kwargs[name] = extras[name]
            elif name in params:
                kwargs[name] = params.pop(name)
            elif p.default is inspect.Parameter.empty and p.kind == p.POSITIONAL_OR_KEYWORD:
                if name == 'vocab':
                    kwargs[name] = None
                else:
                    raise ConfigurationError('missing key "%s" for %s' % (name, sub.__name__))
        if params:
            raise ConfigurationError('Extra parameters passed to %s: %s' % (sub.__name__, params))
        return sub(**kwargs)
