"""Training-step harness around the loss path: not part of the product package."""
