"""amcpy.nn_model -> amcpy_b200.classifier (reference: src/amcpy/nn_model.py:28-75, :88-198, :201-219, :227-267)."""
from amcpy_b200.classifier import AMCClassifier, accuracy_by_snr, load_model, save_model, train_classifier  # noqa: F401
