"""Incremental-state helpers with the reference's key scheme (tell/utils/state.py:10-34):
keys look like 'DynamicConv1dTBC.3.input_buffer'; the decode contract
(filter_incremental_state, decoder_faces_objects.py:175-180) matches on the class-name prefix."""
from collections import defaultdict

INCREMENTAL_STATE_INSTANCE_ID = defaultdict(lambda: 0)


def _full_key(module, key):
    name = module.__class__.__name__
    if not hasattr(module, '_fairseq_instance_id'):
        INCREMENTAL_STATE_INSTANCE_ID[name] += 1
        module._fairseq_instance_id = INCREMENTAL_STATE_INSTANCE_ID[name]
    return '{}.{}.{}'.format(name, module._fairseq_instance_id, key)


def get_incremental_state(module, incremental_state, key):
    full_key = _full_key(module, key)
    if incremental_state is None or full_key not in incremental_state:
        return None
    return incremental_state[full_key]


def set_incremental_state(module, incremental_state, key, value):
    if incremental_state is not None:
        incremental_state[_full_key(module, key)] = value


def eval_str_list(x, type=float):
    """tell/utils/__init__.py eval_str_list: '[1,2]' or list -> list of `type`."""
    if x is None:
        return None
    if isinstance(x, str):
        x = eval(x)
    try:
        return list(map(type, x))
    except TypeError:
        return [type(x)]
