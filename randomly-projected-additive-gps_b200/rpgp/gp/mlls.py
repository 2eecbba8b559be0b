"""ExactMarginalLogLikelihood = [ log N(y | mu, K + sigma^2 I) + sum log priors ] / n   (SURVEY.md Appendix A;
training_routines.py:515,532)."""
from .module import Module


class ExactMarginalLogLikelihood(Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood = likelihood
        self.model = model

    def forward(self, output, target, *params):
        res = self.likelihood(output).log_prob(target)
        for _, prior, closure in self.named_priors():
            res = res + prior.log_prob(closure()).sum()
        return res / target.shape[-1]

    def named_priors(self, prefix=""):
        seen = set()
        for owner in (self.model, self.likelihood):
            for name, prior, closure in owner.named_priors():
                if id(prior) not in seen:
                    seen.add(id(prior))
                    yield name, prior, closure
