"""Trainers with the reference's class names and protocol (reference trainers/*.py): Trainer(sess, config, network),
.train(dataset), .reconstruct(x, dropout=False), .save/.load, .model_dir.  ``sess`` is accepted and ignored."""
