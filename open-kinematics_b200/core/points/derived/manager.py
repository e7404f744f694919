"""Derived-point specification and ordering (reference
core/points/derived/manager.py:31-197).

Unlike the reference, whose derived points are arbitrary Python callables
differentiated by dual numbers, the functions here are *declarative*
(``DerivedFn`` subclasses in ``definitions.py``) so the topology compiler can
lower them to device ops with hand-written JVPs.  They are still callable on a
positions dict, which is how ``initial_state()`` evaluates them once on the host.
"""

from __future__ import annotations

from dataclasses import dataclass


@dataclass
class DerivedPointsSpec:
    functions: dict
    dependencies: dict

    def __post_init__(self) -> None:
        self.validate()

    def validate(self) -> None:
        missing = set(self.functions) - set(self.dependencies)
        if missing:
            raise ValueError(f"Derived points without declared dependencies: {sorted(map(str, missing))}")
        extra = set(self.dependencies) - set(self.functions)
        if extra:
            raise ValueError(f"Dependencies declared for unknown derived points: {sorted(map(str, extra))}")


class DerivedPointsManager:
    """Topologically orders the derived points; raises on cycles."""

    def __init__(self, spec: DerivedPointsSpec):
        self.spec = spec
        self.dependency_graph = spec.dependencies
        self.update_order = self.get_topological_sort()

    def get_topological_sort(self) -> list:
        order: list = []
        state: dict = {}  # 1 = on stack, 2 = done

        def visit(node) -> None:
            mark = state.get(node)
            if mark == 2:
                return
            if mark == 1:
                raise ValueError("Circular dependency detected in derived point definitions.")
            state[node] = 1
            for dep in self.dependency_graph.get(node, ()):
                if dep in self.spec.functions:
                    visit(dep)
            state[node] = 2
            order.append(node)

        for key in self.spec.functions:
            visit(key)
        return order

    def update_in_place(self, positions: dict) -> None:
        for key in self.update_order:
            positions[key] = self.spec.functions[key](positions)
