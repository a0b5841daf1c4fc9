"""Host-side logic that needs no device: the operand-twin registry (tell_b200/twin.py) and the state
copy that lets a captured decode step be reused across generate() calls (Model._copy_tree)."""
import torch


def test_twin_registry_only_serves_the_tensor_it_was_given():
    from tell_b200 import twin
    twin.clear()
    a = torch.randn(8, 16)
    a16 = a.to(torch.bfloat16)
    twin.put(a, a16)
    assert twin.get(a) is a16
    assert twin.get(a.view(8, 16)) is a16            # another tensor object over the same memory
    assert twin.get(a.view(16, 8)) is None           # same pointer, other shape
    assert twin.get(a[:, :8]) is None                # same pointer, other strides
    assert twin.get(a.clone()) is None               # other memory
    a.add_(1.0)                                      # modified after the twin was written
    assert twin.get(a) is None
    b = torch.randn(4, 4)
    twin.put(b, b.to(torch.bfloat16))
    old = twin.enabled
    try:
        twin.enabled = False
        assert twin.get(b) is None
    finally:
        twin.enabled = old
    assert twin.get(b) is not None
    twin.clear()
    assert twin.get(b) is None


def test_twin_registry_keeps_the_memory_alive_and_is_bounded():
    from tell_b200 import twin
    twin.clear()
    a = torch.randn(4, 8)
    ptr = a.data_ptr()
    twin.put(a, a.to(torch.bfloat16))
    del a                                            # the registry still pins the storage:
    c = torch.randn(4, 8)                            # a new tensor cannot land on the registered pointer
    assert c.data_ptr() != ptr
    for _ in range(twin._MAX + 5):                   # never grows without bound
        t = torch.zeros(1, 8)
        twin.put(t, t.to(torch.bfloat16))
    assert len(twin._TW) <= twin._MAX
    twin.clear()


def test_copy_tree_copies_matching_structures_and_rejects_others():
    from tell_b200.models.transformer import _CaptionModelBase
    cp = _CaptionModelBase._copy_tree
    dst = {'a': torch.zeros(2, 3), 'b': [torch.zeros(4), (torch.zeros(1), 7)], 'c': None, 'n': 5, 's': {1, 2}}
    src = {'a': torch.ones(2, 3), 'b': [torch.full((4,), 2.0), (torch.full((1,), 3.0), 7)], 'c': None, 'n': 5,
           's': {3, 4}}
    keep = dst['a']
    assert cp(dst, src)
    assert dst['a'] is keep and torch.equal(dst['a'], src['a'])          # copied INTO the old buffers
    assert torch.equal(dst['b'][0], src['b'][0]) and torch.equal(dst['b'][1][0], src['b'][1][0])
    assert not cp({'a': torch.zeros(2, 3)}, {'a': torch.zeros(3, 2)})     # shape
    assert not cp({'a': torch.zeros(2)}, {'a': torch.zeros(2, dtype=torch.float64)})   # dtype
    assert not cp({'a': torch.zeros(2)}, {'b': torch.zeros(2)})           # keys
    assert not cp([torch.zeros(2)], [torch.zeros(2), torch.zeros(2)])     # length
    assert not cp({'a': torch.zeros(2)}, {'a': None})                     # tensor vs not
    assert not cp({'n': 5}, {'n': 6})                                     # scalar state that differs
