"""Form tags (thermca/fem/fe_system.py:29)."""


class _Form:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return f"<form {self.name}>"


laplace = _Form("laplace")
mass = _Form("mass")
unit_load = _Form("unit_load")
