"""Drop-in mirror of the reference package `u2pl` (only the names train_semi.py / train_sup.py
import: SURVEY.md section 8b).  `u2pl_b200.install()` puts this directory's parent on sys.path so
`from u2pl.utils.loss_helper import compute_unsupervised_loss` resolves here."""
