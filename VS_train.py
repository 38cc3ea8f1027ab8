"""Train the 2.5D attention U-Net (same recipe and call order as the reference's VS_train.py:15-51).
usage: python VS_train.py [--debug] [--dataset T1|T2] [--results_folder_name NAME] [--synthetic] [--device cpu|cuda:0]"""
import argparse

from params.VSparams import VSparams
from vs_seg_b200.dataio import set_determinism

parser = argparse.ArgumentParser(description="Train the model")

# initialize parameters
p = VSparams(parser)

# create folders
p.create_results_folders()

# set up logger
logger = p.set_up_logger("training_log.txt")

# log parameters
p.log_parameters()

# load paths to data sets
train_files, val_files, test_files = p.load_T1_or_T2_data()

# define the transforms
train_transforms, val_transforms, test_transforms = p.get_transforms()

# Set deterministic training for reproducibility
set_determinism(seed=0)

# check transforms
p.check_transforms_on_first_validation_image_and_label(val_files, val_transforms)

# cache and load data
train_loader = p.cache_transformed_train_data(train_files, train_transforms)
val_loader = p.cache_transformed_val_data(val_files, val_transforms)

# set model, loss function and optimizer
model = p.set_and_get_model()
loss_function = p.set_and_get_loss_function()
optimizer = p.set_and_get_optimizer(model)

# run training algorithm
epoch_loss_values, metric_values = p.run_training_algorithm(model, loss_function, optimizer, train_loader, val_loader)

# plot loss and mean dice
p.plot_loss_curve_and_mean_dice(epoch_loss_values, metric_values)
