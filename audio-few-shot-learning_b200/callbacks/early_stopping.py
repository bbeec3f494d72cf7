"""Early stopping on validation accuracy (mirror of callbacks/early_stopping.py:15-70).

Same constructor, same call protocol ``stopper(val_accuracy, model, epoch)``, same checkpoint file
(``torch.save(model.state_dict(), path)``): the counter resets on improvement beyond ``delta``, a message is
traced once the counter reaches 80 % of the patience, ``early_stop`` is raised at the patience.
"""
import numpy as np
from torch import save

GREEN, RED, ENDC = "\033[92m", "\033[91m", "\033[0m"


class EarlyStopping:
    def __init__(self, patience=7, verbose=False, delta=0, path="checkpoint.pt", trace_func=print):
        self.patience = patience
        self.verbose = verbose
        self.counter = 0
        self.best_score = None
        self.early_stop = False
        self.val_accuracy_max = -np.inf
        self.delta = delta
        self.path = path
        self.trace_func = trace_func

    def __call__(self, val_accuracy, model, epoch):
        score = val_accuracy
        if self.best_score is None:
            self.best_score = score
            self.save_checkpoint(val_accuracy, model, epoch)
        elif score < self.best_score + self.delta:
            self.counter += 1
            if self.counter >= int(0.8 * self.patience):
                self.trace_func(f"Epoch: {epoch}. EarlyStopping counter: {self.counter} out of {self.patience}")
            if self.counter >= self.patience:
                self.early_stop = True
        else:
            self.best_score = score
            self.save_checkpoint(val_accuracy, model, epoch)
            self.counter = 0

    def save_checkpoint(self, val_accuracy, model, epoch):
        """Saves the state dict when the validation accuracy improved."""
        if self.verbose:
            increase = (val_accuracy - self.val_accuracy_max) / self.val_accuracy_max * 100 if self.val_accuracy_max > 0 else 0
            colour = GREEN if increase > 0 else RED
            self.trace_func(f"Epoch: {epoch}. Validation accuracy increased ({self.val_accuracy_max:.6f} --> "
                            f"{val_accuracy:.6f}), {colour}({increase:.2f}%){ENDC} Saving model ...")
        save(model.state_dict(), self.path)
        self.val_accuracy_max = val_accuracy
