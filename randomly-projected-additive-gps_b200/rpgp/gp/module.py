"""Module base: raw parameters + constraints + priors + `initialize(**values)` (the gpytorch.Module surface the
reference uses: kernel.initialize(lengthscale=...), likelihood.noise = ..., test.py:126-134)."""
import torch


class Module(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self._priors = {}

    def register_constraint(self, param_name, constraint):
        self.add_module(param_name + "_constraint", constraint)

    def constraint_for(self, param_name):
        return getattr(self, param_name + "_constraint")

    def register_prior(self, name, prior, closure):
        """closure(module) -> the value the prior is evaluated on"""
        self.add_module(name, prior)
        self._priors[name] = closure

    def named_priors(self, prefix=""):
        for mname, module in self.named_modules(prefix=prefix):
            for pname, closure in getattr(module, "_priors", {}).items():
                yield (mname + "." if mname else "") + pname, getattr(module, pname), (lambda c=closure, m=module: c(m))

    def initialize(self, **kwargs):
        """Set parameters by name.  Names of constrained values (lengthscale, outputscale, noise, ...) go through their
        property setters; raw names and plain parameters are copied in place.  'a.b' paths reach sub-modules."""
        for name, val in kwargs.items():
            if "." in name:
                head, tail = name.split(".", 1)
                getattr(self, head).initialize(**{tail: val})
                continue
            if not hasattr(self, name):
                raise AttributeError("Unknown parameter {p} for {c}".format(p=name, c=self.__class__.__name__))
            if name in self._parameters:
                param = self._parameters[name]
                with torch.no_grad():
                    v = torch.as_tensor(val, dtype=param.dtype, device=param.device)
                    if v.numel() == param.numel():
                        param.copy_(v.reshape(param.shape))
                    elif v.numel() == 1:
                        param.copy_(v.reshape(()).expand_as(param))
                    else:
                        raise ValueError("cannot initialise %s of shape %s with %d values"
                                         % (name, tuple(param.shape), v.numel()))
            else:
                setattr(self, name, val)
        return self

    def _set_constrained(self, raw_name, value):
        raw = self._parameters[raw_name]
        value = torch.as_tensor(value, dtype=raw.dtype, device=raw.device)
        self.initialize(**{raw_name: self.constraint_for(raw_name).inverse_transform(value)})
