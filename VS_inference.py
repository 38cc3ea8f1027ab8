"""Sliding-window inference on the test split with the best checkpoint (same recipe and call order as the
reference's VS_inference.py:15-42).  Under `torchrun --nproc-per-node N` the windows of every volume are
sharded over the N GPUs (one NCCL reduce per volume).
usage: python VS_inference.py [--debug] [--dataset T1|T2] [--results_folder_name NAME] [--device cpu|cuda:0]"""
import argparse

from params.VSparams import VSparams
from vs_seg_b200.dataio import set_determinism

parser = argparse.ArgumentParser(description="Run inference with the trained model")

# initialize parameters
p = VSparams(parser)

# set up logger
logger = p.set_up_logger("test_log.txt")

# log parameters
p.log_parameters()

# load paths to data sets
train_files, val_files, test_files = p.load_T1_or_T2_data()

# define the transforms
train_transforms, val_transforms, test_transforms = p.get_transforms()

# Set deterministic training for reproducibility
set_determinism(seed=0)

# cache and load validation data
test_loader = p.cache_transformed_test_data(test_files, test_transforms)

# set model
model = p.set_and_get_model()

# load the trained model and set it into evaluation mode
model = p.load_trained_state_of_model(model)

# run inference and create figures in figures folder
p.run_inference(model, test_loader)
