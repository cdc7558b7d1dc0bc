"""Analysis wrappers registered on Model (reference fava/analysis/*.py): each forwards to the loaded
mesh, wrapped by the `timer` decorator exactly like `@Model.register_analysis(use_timer=True)`."""

from fava_b200.model import Model


@Model.register_analysis(use_timer=True)
def reynolds_stress(self, *args, **kwargs):
    return self.mesh.reynolds_stress(*args, **kwargs)


@Model.register_analysis(use_timer=True)
def favre_stress(self, *args, **kwargs):
    return self.mesh.favre_stress(*args, **kwargs)


@Model.register_analysis(use_timer=True)
def kinetic_energy_spectra(self, *args, **kwargs):
    return self.mesh.kinetic_energy_spectra(*args, **kwargs)


@Model.register_analysis(use_timer=True)
def slice_average(self, *args, **kwargs):
    return self.mesh.slice_average(*args, **kwargs)


@Model.register_analysis(use_timer=True)
def slice_integral(self, *args, **kwargs):
    return self.mesh.slice_integral(*args, **kwargs)


@Model.register_analysis(use_timer=True)
def from_amr(self, *args, **kwargs):
    return self.mesh.from_amr(*args, **kwargs)


@Model.register_analysis(use_timer=True)
def fractal_dimension(self, *args, **kwargs):
    return self.mesh.fractal_dimension(*args, **kwargs)


@Model.register_analysis(use_timer=True)
def structure_functions(self, *args, **kwargs):
    return self.mesh.structure_functions(*args, **kwargs)
