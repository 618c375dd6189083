"""Adversarial domain-adaptation pieces of the reference's src/DA package that are on the hot path:
grl.WarmStartGradientReverseLayer and cdan_frame.ConditionalDomainAdversarialLoss (clip-level BCE form)."""
