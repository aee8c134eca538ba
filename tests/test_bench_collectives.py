"""bench.py under torchrun: a collective that only some ranks join hangs until the NCCL watchdog aborts the
communicator (this is what killed every N > 1 run of round 1: an all-reduce inside ``if rank == 0:``).

Static guard (CPU): no collective call may sit lexically inside an ``if`` / ``while`` / conditional expression
whose test reads ``rank``, and no function that issues collectives may be called from such a place either."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# names that are (or wrap) collectives in bench.py / trackmpnn_b200.parallel
COLLECTIVE_CALLS = {'reduce_', 'barrier', 'all_reduce', 'all_gather', 'all_gather_into_tensor', 'broadcast', 'reduce_scatter',
                    'allreduce_gradients', 'allreduce_flat', 'timed_max', 'reduce_sum', 'init_process_group',
                    'destroy_process_group'}


def _call_name(node):
    f = node.func
    if isinstance(f, ast.Name):
        return f.id
    if isinstance(f, ast.Attribute):
        return f.attr
    return None


def _reads_rank(test):
    return any(isinstance(n, ast.Name) and n.id in ('rank', 'local', 'local_rank') for n in ast.walk(test))


def _collective_functions(tree):
    """Functions of the module whose body (transitively) issues a collective."""
    funcs = {n.name: n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)}
    direct = {name for name, fn in funcs.items()
              if any(isinstance(c, ast.Call) and _call_name(c) in COLLECTIVE_CALLS for c in ast.walk(fn))}
    changed = True
    while changed:
        changed = False
        for name, fn in funcs.items():
            if name in direct:
                continue
            if any(isinstance(c, ast.Call) and _call_name(c) in direct for c in ast.walk(fn)):
                direct.add(name)
                changed = True
    direct.discard('main')
    return direct


def _violations(path):
    with open(path) as f:
        tree = ast.parse(f.read())
    guarded = COLLECTIVE_CALLS | _collective_functions(tree)
    bad = []

    def visit(node, under_rank):
        if isinstance(node, (ast.If, ast.While, ast.IfExp)):
            inner = under_rank or _reads_rank(node.test)
            body = node.body if isinstance(node.body, list) else [node.body]
            orelse = node.orelse if isinstance(node.orelse, list) else [node.orelse]
            for ch in body + orelse:
                visit(ch, inner)
            visit(node.test, under_rank)
            return
        if isinstance(node, ast.Call) and under_rank and _call_name(node) in guarded:
            bad.append((_call_name(node), node.lineno))
        for ch in ast.iter_child_nodes(node):
            visit(ch, under_rank)

    visit(tree, False)
    return bad


def test_no_collective_under_a_rank_test_in_bench():
    assert _violations(os.path.join(ROOT, 'bench.py')) == []


def test_guard_catches_the_round1_bug(tmp_path):
    p = tmp_path / 'b.py'
    p.write_text('def main():\n    x = reduce_(1, 2)\n    if rank == 0:\n        out = {"a": reduce_(3, 4)}\n')
    assert _violations(str(p)) == [('reduce_', 4)]
    p.write_text('def leg(a):\n    barrier()\n\ndef main():\n    if world > 1 and rank != 1:\n        leg(1)\n')
    assert _violations(str(p)) == [('leg', 6)]
    p.write_text('def main():\n    if world > 1:\n        barrier()\n    y = reduce_(1, 2) if world > 1 else 0\n')
    assert _violations(str(p)) == []
