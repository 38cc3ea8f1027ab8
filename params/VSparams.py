"""VSparams: configuration + orchestration object of the VS_Seg entry points, with the reference's
method names, flags, folder layout and call order (/root/reference/params/VSparams.py:37-619), driving
the B200-native hot path:

  * model / loss / inferer are this repo's drop-ins (`UNet2d5_spvPA`, `Dice_spvPA`,
    `vs_seg_b200.sliding_window.sliding_window_inference`);
  * the data plane is the MONAI-free stand-in of `vs_seg_b200.dataio` (the image has no MONAI, nibabel
    or matplotlib); figures degrade to PNGs written with PIL, or are skipped;
  * additive flags only: `--device` (default cuda:0, falls back to cpu when no GPU is visible — the CPU
    path is BASELINE config 0 plumbing), `--synthetic` (write synthetic NIfTI cases for the split when the
    dataset is absent), `--num_epochs`, `--data_root`.
  * under `torchrun` (WORLD_SIZE > 1) inference shards the sliding-window windows over the ranks.
"""
import csv
import logging
import os
from time import perf_counter, strftime

import numpy as np
import torch
from torch.utils.data import DataLoader

from vs_seg_b200 import dataio
from vs_seg_b200.compat import Norm
from vs_seg_b200.dataio import (AddChanneld, Compose, LoadNiftid, NiftiSaver, NormalizeIntensityd, Orientationd,
                                RandFlipd, RandSpatialCropd, SpatialPadd, ToTensord)
from vs_seg_b200.sliding_window import sliding_window_inference

from .losses.dice_spvPA import Dice_spvPA, DiceLoss
from .networks.nets.unet2d5_spvPA import UNet2d5_spvPA


def _save_png(path, panels):
    """Side-by-side grayscale panels (2-D arrays) as one PNG; silently skipped without PIL."""
    try:
        from PIL import Image
    except ImportError:
        return False
    tiles = []
    for p in panels:
        p = np.asarray(p, dtype=np.float32)
        lo, hi = float(p.min()), float(p.max())
        tiles.append(((p - lo) / (hi - lo + 1e-12) * 255).astype(np.uint8).T)
    h = max(t.shape[0] for t in tiles)
    canvas = np.concatenate([np.pad(t, ((0, h - t.shape[0]), (0, 4))) for t in tiles], axis=1)
    Image.fromarray(canvas).save(path)
    return True


def _plot_png(path, panels, size=(600, 400)):
    """Line plots side by side with PIL (the image has no matplotlib): panels = [(title, xs, ys, xlabel)]."""
    try:
        from PIL import Image, ImageDraw
    except ImportError:
        return False
    W, H = size
    img = Image.new("RGB", (W * len(panels), H), "white")
    d = ImageDraw.Draw(img)
    for k, (title, xs, ys, xlabel) in enumerate(panels):
        x0, y0, x1, y1 = k * W + 60, 40, (k + 1) * W - 20, H - 50
        d.rectangle([x0, y0, x1, y1], outline="black")
        d.text((x0 + 5, 12), title, fill="black")
        d.text(((x0 + x1) // 2 - 15, H - 25), xlabel, fill="black")
        if len(xs):
            lo, hi = float(min(ys)), float(max(ys))
            hi = hi if hi > lo else lo + 1.0
            xl, xh = float(min(xs)), float(max(xs))
            xh = xh if xh > xl else xl + 1.0
            pts = [(x0 + (x - xl) / (xh - xl) * (x1 - x0), y1 - (y - lo) / (hi - lo) * (y1 - y0)) for x, y in zip(xs, ys)]
            if len(pts) > 1:
                d.line(pts, fill=(31, 119, 180), width=2)
            for px, py in pts:
                d.ellipse([px - 2, py - 2, px + 2, py + 2], fill=(31, 119, 180))
            d.text((k * W + 5, y0 - 5), f"{hi:.4g}", fill="black")
            d.text((k * W + 5, y1 - 5), f"{lo:.4g}", fill="black")
            d.text((x0, y1 + 5), f"{xl:g}", fill="black")
            d.text((x1 - 25, y1 + 5), f"{xh:g}", fill="black")
    img.save(path)
    return True


def _hist_png(path, values, bins, size=(600, 400)):
    """Histogram bar chart with PIL (reference: plt.hist(dice_scores, bins=np.arange(0, 1.01, 0.01)))."""
    try:
        from PIL import Image, ImageDraw
    except ImportError:
        return False
    counts, edges = np.histogram(np.asarray(values, dtype=np.float64), bins=bins)
    W, H = size
    img = Image.new("RGB", (W, H), "white")
    d = ImageDraw.Draw(img)
    x0, y0, x1, y1 = 50, 30, W - 20, H - 40
    d.rectangle([x0, y0, x1, y1], outline="black")
    top = max(int(counts.max()), 1)
    bw = (x1 - x0) / len(counts)
    for i, c in enumerate(counts):
        if c:
            d.rectangle([x0 + i * bw, y1 - c / top * (y1 - y0), x0 + (i + 1) * bw, y1], fill=(31, 119, 180))
    d.text((5, y0 - 5), str(top), fill="black")
    d.text((x0, y1 + 5), f"{edges[0]:g}", fill="black")
    d.text((x1 - 20, y1 + 5), f"{edges[-1]:g}", fill="black")
    img.save(path)
    return True


class VSparams:
    def __init__(self, parser):
        parser.add_argument("--debug", dest="debug", action="store_true", help="activate debugging mode")
        parser.set_defaults(debug=False)
        parser.add_argument("--split", type=str, default="./params/split_TCIA.csv",
                            help="path to CSV file that defines training, validation and test datasets")
        parser.add_argument("--dataset", type=str, default="T1", help='(string) use "T1" or "T2" to select dataset')
        parser.add_argument("--train_batch_size", type=int, default=1, help="batch size of the forward pass")
        parser.add_argument("--initial_learning_rate", type=float, default=1e-4, help="learning rate at first epoch")
        parser.add_argument("--no_attention", dest="attention", action="store_false",
                            help="disables the attention module in the network and the attention map weighting in "
                                 "the loss function")
        parser.set_defaults(attention=True)
        parser.add_argument("--no_hardness", dest="hardness", action="store_false",
                            help="disables the hardness weighting in the loss function")
        parser.set_defaults(hardness=True)
        parser.add_argument("--results_folder_name", type=str, default="temp" + strftime("%Y%m%d%H%M%S"),
                            help="name of results folder")
        # additive flags (not in the reference)
        parser.add_argument("--device", type=str, default=None, help='torch device (default "cuda:0", cpu if no GPU)')
        parser.add_argument("--synthetic", action="store_true",
                            help="write synthetic NIfTI cases for the split if the dataset is missing")
        parser.add_argument("--num_epochs", type=int, default=None, help="override the number of epochs")
        parser.add_argument("--data_root", type=str, default="./data/VS_defaced/", help="path to the data set")

        args = parser.parse_args()

        self.debug = args.debug
        self.dataset = args.dataset
        self.data_root = args.data_root
        self.split_csv = args.split
        if self.debug:
            self.split_csv = "./params/split_debug.csv"
        self.pad_crop_shape = [384, 384, 64]
        if self.debug:
            self.pad_crop_shape = [128, 128, 32]
        self.pad_crop_shape_test = [384, 384, 64]
        if self.debug:
            self.pad_crop_shape_test = [128, 128, 32]
        self.num_workers = 4
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        # the reference hard-codes cuda:0 (VSparams.py:83); under torchrun every rank drives its own GPU
        default_gpu = f"cuda:{self.local_rank}" if self.world_size > 1 else "cuda:0"
        self.torch_device_arg = args.device or (default_gpu if torch.cuda.is_available() else "cpu")
        self.train_batch_size = args.train_batch_size
        self.initial_learning_rate = args.initial_learning_rate
        self.epochs_with_const_lr = 100
        if self.debug:
            self.epochs_with_const_lr = 3
        self.lr_divisor = 2.0
        self.weight_decay = 1e-7
        self.num_epochs = 300
        if self.debug:
            self.num_epochs = 10
        if args.num_epochs is not None:
            self.num_epochs = args.num_epochs
        self.val_interval = 2  # determines how frequently validation is performed during training
        self.model = "UNet2d5_spvPA"
        self.sliding_window_inferer_roi_size = [384, 384, 64]
        if self.debug:
            self.sliding_window_inferer_roi_size = [128, 128, 32]
        self.attention = args.attention
        self.hardness = args.hardness
        self.export_inferred_segmentations = True
        self.synthetic = args.synthetic
        self.synthetic_shape = (64, 64, 64)

        # paths
        self.results_folder_path = os.path.join(self.data_root, "results", args.results_folder_name)
        if self.debug:
            self.results_folder_path = os.path.join(self.data_root, "results", "debug")
        self.logs_path = os.path.join(self.results_folder_path, "logs")
        self.model_path = os.path.join(self.results_folder_path, "model")
        self.figures_path = os.path.join(self.results_folder_path, "figures")

        self.device = torch.device(self.torch_device_arg)
        if self.device.type == "cuda":
            torch.cuda.set_device(self.device)
        if self.world_size > 1:
            # one process group for the whole run, created before any rank touches the data set
            import torch.distributed as dist
            if not dist.is_initialized():
                dist.init_process_group("nccl" if self.device.type == "cuda" else "gloo")

    def create_results_folders(self):
        for path in (self.logs_path, self.model_path, self.figures_path):
            if not os.path.exists(path):
                os.makedirs(path, exist_ok=True)
                os.chmod(path, 0o777)

    def set_up_logger(self, log_file_name):
        os.makedirs(self.logs_path, exist_ok=True)
        self.logger = logging.getLogger()
        if self.rank != 0:   # replicas log next to rank 0's file instead of truncating it
            stem, ext = os.path.splitext(log_file_name)
            log_file_name = f"{stem}_rank{self.rank}{ext}"
        file_handler = logging.FileHandler(os.path.join(self.logs_path, log_file_name), mode="w")
        console_handler = logging.StreamHandler()
        self.logger.addHandler(file_handler)
        self.logger.addHandler(console_handler)
        formatter = logging.Formatter("%(asctime)s %(levelname)s        %(message)s")
        file_handler.setFormatter(formatter)
        console_handler.setFormatter(formatter)
        self.logger.setLevel(logging.INFO)
        self.logger.info("Created " + log_file_name)
        return self.logger

    def log_parameters(self):
        logger = self.logger
        logger.info("-" * 10)
        logger.info("Parameters: ")
        for name in ("dataset", "data_root", "split_csv", "pad_crop_shape", "pad_crop_shape_test", "num_workers",
                     "torch_device_arg", "train_batch_size", "initial_learning_rate", "epochs_with_const_lr",
                     "lr_divisor", "weight_decay", "num_epochs", "val_interval", "model",
                     "sliding_window_inferer_roi_size", "attention", "hardness", "results_folder_path",
                     "export_inferred_segmentations"):
            logger.info("{:<34s} {}".format(name + " =", getattr(self, name)))
        logger.info("-" * 10)

    def load_T1_or_T2_data(self):
        logger = self.logger
        train_files, val_files, test_files = [], [], []
        if self.synthetic and self.rank == 0:
            missing = True
            with open(self.split_csv) as f:
                first = next(csv.reader(f))[0]
            probe = "vs_gk_t1_refT1.nii.gz" if self.dataset == "T1" else "vs_gk_t2_refT2.nii.gz"
            missing = not os.path.isfile(os.path.join(self.data_root, "input_data", first, probe))
            if missing:
                logger.info("Writing synthetic cases for the split (no dataset present)...")
                dataio.make_synthetic_dataset(self.data_root, self.split_csv, self.dataset, self.synthetic_shape)
        if self.synthetic and self.world_size > 1:
            import torch.distributed as dist
            dist.barrier()   # the other ranks wait for rank 0's files
        with open(self.split_csv) as csvfile:
            for row in csv.reader(csvfile):
                if not row:
                    continue
                if self.dataset == "T1":
                    image_name = os.path.join(self.data_root, "input_data", row[0], "vs_gk_t1_refT1.nii.gz")
                    label_name = os.path.join(self.data_root, "input_data", row[0], "vs_gk_seg_refT1.nii.gz")
                elif self.dataset == "T2":
                    image_name = os.path.join(self.data_root, "input_data", row[0], "vs_gk_t2_refT2.nii.gz")
                    label_name = os.path.join(self.data_root, "input_data", row[0], "vs_gk_seg_refT2.nii.gz")
                else:
                    raise ValueError('--dataset must be "T1" or "T2"')
                item = {"image": image_name, "label": label_name}
                {"training": train_files, "validation": val_files, "test": test_files}.get(row[1], []).append(item)

        for file_dict in train_files + val_files + test_files:
            assert os.path.isfile(file_dict["image"]), f" {file_dict['image']} is not a file"
            assert os.path.isfile(file_dict["label"]), f" {file_dict['label']} is not a file"

        logger.info("Number of images in training set   = {}".format(len(train_files)))
        logger.info("Number of images in validation set = {}".format(len(val_files)))
        logger.info("Number of images in test set       = {}".format(len(test_files)))
        logger.info("training set   = {}".format(train_files))
        logger.info("validation set = {}".format(val_files))
        logger.info("test set       = {}".format(test_files))
        return train_files, val_files, test_files

    def get_transforms(self):
        self.logger.info("Getting transforms...")
        keys = ["image", "label"]
        head = lambda: [LoadNiftid(keys=keys), AddChanneld(keys=keys), Orientationd(keys=keys, axcodes="RAS"),  # noqa: E731
                        NormalizeIntensityd(keys=["image"])]
        train_transforms = Compose(head() + [
            SpatialPadd(keys=keys, spatial_size=self.pad_crop_shape),
            RandFlipd(keys=keys, prob=0.5, spatial_axis=0),
            RandSpatialCropd(keys=keys, roi_size=self.pad_crop_shape, random_center=True, random_size=False),
            ToTensord(keys=keys)])
        val_transforms = Compose(head() + [
            SpatialPadd(keys=keys, spatial_size=self.pad_crop_shape),
            RandSpatialCropd(keys=keys, roi_size=self.pad_crop_shape, random_center=True, random_size=False),
            ToTensord(keys=keys)])
        test_transforms = Compose(head() + [ToTensord(keys=keys)])
        return train_transforms, val_transforms, test_transforms

    @staticmethod
    def get_center_of_mass_slice(label):
        """Through-plane slice closest to the label's centre of mass (middle slice for an empty label)."""
        label = np.asarray(label)
        num_slices = label.shape[2]
        slice_masses = label.reshape(-1, num_slices).sum(0).astype(np.float64)
        if slice_masses.sum() == 0:
            slice_weights = np.ones(num_slices) / num_slices
        else:
            slice_weights = slice_masses / slice_masses.sum()
        return int(np.round((slice_weights * np.arange(num_slices)).sum()))

    def check_transforms_on_first_validation_image_and_label(self, val_files, val_transforms):
        logger = self.logger
        check_ds = dataio.ArrayDataset(data=val_files, transform=val_transforms)
        check_data = next(iter(DataLoader(check_ds, batch_size=1)))
        image, label = check_data["image"][0][0], check_data["label"][0][0]
        logger.info("-" * 10)
        logger.info("Check the transforms on the first validation set image and label")
        logger.info("Length of check_data = {}".format(len(check_data)))
        logger.info("check_data['image'].shape = {}".format(check_data["image"].shape))
        logger.info("Validation image shape = {}".format(image.shape))
        logger.info("Validation label shape = {}".format(label.shape))
        slice_idx = self.get_center_of_mass_slice(label)
        logger.info("-" * 10)
        logger.info("Plot one slice of the image and the label")
        logger.info("image shape: {}, label shape: {}, slice = {}".format(image.shape, label.shape, slice_idx))
        os.makedirs(self.figures_path, exist_ok=True)
        _save_png(os.path.join(self.figures_path, "check_validation_image_and_label.png"),
                  [image[:, :, slice_idx], label[:, :, slice_idx]])

    @staticmethod
    def worker_init_fn(worker_id):
        worker_info = torch.utils.data.get_worker_info()
        worker_info.dataset.transform.set_random_state(worker_info.seed % (2 ** 32))

    def cache_transformed_train_data(self, train_files, train_transforms):
        self.logger.info("Caching training data set...")
        if self.world_size > 1:   # data-parallel: every rank trains on its own shard (DistributedSampler semantics)
            from vs_seg_b200.ddp import shard_list
            train_files = shard_list(train_files, self.rank, self.world_size)
        train_ds = dataio.CacheDataset(data=train_files, transform=train_transforms, cache_rate=1.0,
                                       num_workers=self.num_workers)
        return DataLoader(train_ds, batch_size=self.train_batch_size, shuffle=True, num_workers=self.num_workers,
                          collate_fn=dataio.list_data_collate, worker_init_fn=self.worker_init_fn)

    def cache_transformed_val_data(self, val_files, val_transforms):
        self.logger.info("Caching validation data set...")
        val_ds = dataio.CacheDataset(data=val_files, transform=val_transforms, cache_rate=1.0,
                                     num_workers=self.num_workers)
        return DataLoader(val_ds, batch_size=1, num_workers=self.num_workers)

    def cache_transformed_test_data(self, test_files, test_transforms):
        self.logger.info("Caching test data set...")
        test_ds = dataio.CacheDataset(data=test_files, transform=test_transforms, cache_rate=1.0,
                                      num_workers=self.num_workers)
        return DataLoader(test_ds, batch_size=1, num_workers=self.num_workers)

    def set_and_get_model(self):
        logger = self.logger
        logger.info("Setting up the model type...")
        if self.model != "UNet2d5_spvPA":
            raise ValueError(f"unknown model {self.model}")
        model = UNet2d5_spvPA(
            dimensions=3,
            in_channels=1,
            out_channels=2,
            channels=(16, 32, 48, 64, 80, 96),
            strides=((2, 2, 1), (2, 2, 1), (2, 2, 2), (2, 2, 2), (2, 2, 2)),
            kernel_sizes=((3, 3, 1), (3, 3, 1), (3, 3, 3), (3, 3, 3), (3, 3, 3), (3, 3, 3)),
            sample_kernel_sizes=((3, 3, 1), (3, 3, 1), (3, 3, 3), (3, 3, 3), (3, 3, 3)),
            num_res_units=2,
            norm=Norm.BATCH,
            dropout=0.1,
            attention_module=self.attention,
        ).to(self.device)
        return model

    def set_and_get_loss_function(self):
        self.logger.info("Setting up the loss function...")
        return Dice_spvPA(to_onehot_y=True, softmax=True, supervised_attention=self.attention,
                          hardness_weighting=self.hardness)

    def set_and_get_optimizer(self, model):
        self.logger.info("Setting up the optimizer...")
        # same interface as the reference's torch.optim.Adam (VSparams.py:388-391); CUDA parameters are stepped
        # by one fused native launch over a flat buffer, CPU parameters by torch.optim.Adam itself
        from vs_seg_b200.optim import FusedAdam
        return FusedAdam(model.parameters(), lr=self.initial_learning_rate, weight_decay=self.weight_decay)

    def compute_dice_score(self, predicted_probabilities, label):
        """Hard foreground Dice of the argmax segmentation (reference VSparams.py:393-408).  CUDA tensors: one
        native launch (argmax + Dice sums, vsseg_sw_finalize), result stays on the device - no host sync."""
        if predicted_probabilities.is_cuda:
            from vs_seg_b200.sliding_window import hard_dice
            return hard_dice(predicted_probabilities, label).mean().float().reshape(1, 1)
        from vs_seg_b200.compat import one_hot
        n_classes = predicted_probabilities.shape[1]
        y_pred = torch.argmax(predicted_probabilities, dim=1, keepdim=True)
        y_pred = one_hot(y_pred, n_classes)
        loss = DiceLoss(include_background=False, to_onehot_y=True, softmax=False, reduction="mean")(y_pred, label)
        return torch.tensor([[1 - loss]], device=self.device)

    def run_training_algorithm(self, model, loss_function, optimizer, train_loader, val_loader):
        logger = self.logger
        logger.info("Running the training loop...")
        tb_writer = None
        if self.rank == 0:
            try:
                from torch.utils.tensorboard import SummaryWriter
                tb_writer = SummaryWriter()
            except Exception:  # tensorboard is optional plumbing
                tb_writer = None

        epochs_with_const_lr = self.epochs_with_const_lr
        val_interval = self.val_interval
        best_metric, best_metric_epoch = -1, -1
        epoch_loss_values, metric_values = list(), list()
        num_epochs = self.num_epochs
        # data-parallel training under torchrun: weights of rank 0 everywhere, one all-reduce of the flat
        # gradient per step (vs_seg_b200.ddp); a single process is the reference's own loop
        from vs_seg_b200 import ddp
        if self.world_size > 1:
            ddp.broadcast_module_state(model)
        reducer = ddp.GradReducer(model, optimizer)
        # CUDA: the loop body below (zero_grad .. optimizer.step, reference VSparams.py:457-462) is captured once per batch
        # shape in a CUDA graph and replayed - the eager step is host-bound (~420 launches); VSSEG_GRAPH_STEP=0 disables
        graphed_step = None
        if self.device.type == "cuda" and os.environ.get("VSSEG_GRAPH_STEP", "1") != "0" and hasattr(optimizer, "pre_replay"):
            from vs_seg_b200.training import GraphedTrainStep
            graphed_step = GraphedTrainStep(model, loss_function, optimizer, reducer)
        start = perf_counter()
        for epoch in range(num_epochs):
            logger.info("-" * 10)
            logger.info("Epoch {}/{}".format(epoch + 1, num_epochs))
            if epoch == val_interval:
                stop = perf_counter()
                logger.info(("Average duration of first {0:.0f} epochs = {1:.2f} s. "
                             "Expected total training time = {2:.2f} h").format(
                    val_interval, (stop - start) / val_interval, (stop - start) * num_epochs / val_interval / 3600))
            model.train()
            epoch_loss, step = 0, 0
            for batch_data in train_loader:
                step += 1
                inputs, labels = batch_data["image"].to(self.device), batch_data["label"].to(self.device)
                if graphed_step is not None:
                    loss = graphed_step(inputs, labels)
                else:
                    optimizer.zero_grad()
                    outputs = model(inputs)
                    loss = loss_function(outputs, labels)
                    loss.backward()
                    reducer.reduce()
                    optimizer.step()
                # the loss stays on the device (no host sync per step); it is read once per epoch / log line
                epoch_loss = epoch_loss + loss.detach()
                if epoch == 0:
                    logger.info("{}/{}, train_loss: {:.4f}".format(step, len(train_loader) // train_loader.batch_size,
                                                                   loss.item()))
            epoch_loss = float(epoch_loss) / max(step, 1)
            epoch_loss_values.append(epoch_loss)
            logger.info("epoch {} average loss: {:.4f}".format(epoch + 1, epoch_loss))

            if (epoch + 1) % val_interval == 0:
                model.eval()
                with torch.no_grad():
                    metric_sum, metric_count, epoch_loss_val, step = 0.0, 0, 0, 0
                    for val_data in val_loader:
                        step += 1
                        val_inputs, val_labels = val_data["image"].to(self.device), val_data["label"].to(self.device)
                        val_outputs = model(val_inputs)
                        dice_score = self.compute_dice_score(val_outputs[0], val_labels)
                        loss = loss_function(val_outputs, val_labels)
                        # the reference accumulates these twice per image (VSparams.py:490-496), which doubles the
                        # logged validation loss and leaves the metric ratio unchanged; kept for log compatibility.
                        # The sums stay on the device: one host read per validation pass instead of three per image
                        for _ in range(2):
                            metric_count += len(dice_score)
                            metric_sum = metric_sum + dice_score.sum()
                            epoch_loss_val = epoch_loss_val + loss.detach()
                    metric_sum, epoch_loss_val = float(metric_sum), float(epoch_loss_val)
                    metric = metric_sum / metric_count
                    metric_values.append(metric)
                    epoch_loss_val /= step
                    if tb_writer is not None:
                        tb_writer.add_scalars("Loss Train/Val", {"train": epoch_loss, "val": epoch_loss_val}, epoch)
                        tb_writer.add_scalar("Dice Score Val", metric, epoch)
                    if metric > best_metric:
                        best_metric, best_metric_epoch = metric, epoch + 1
                        if self.rank == 0:   # the replicas hold identical weights
                            torch.save(model.state_dict(), os.path.join(self.model_path, "best_metric_model.pth"))
                        logger.info("saved new best metric model")
                    logger.info("current epoch {} current mean dice: {:.4f} best mean dice: {:.4f} at epoch {}".format(
                        epoch + 1, metric, best_metric, best_metric_epoch))

            if (epoch + 1) % epochs_with_const_lr == 0:
                for param_group in optimizer.param_groups:
                    param_group["lr"] = param_group["lr"] / self.lr_divisor
                    logger.info("Dividing learning rate by {}. New learning rate is: lr = {}".format(
                        self.lr_divisor, param_group["lr"]))

        logger.info("Train completed, best_metric: {:.4f}  at epoch: {}".format(best_metric, best_metric_epoch))
        if self.rank == 0:
            torch.save(model.state_dict(), os.path.join(self.model_path, "last_epoch_model.pth"))
        logger.info(f'Saved model of the last epoch at: {os.path.join(self.model_path, "last_epoch_model.pth")}')
        return epoch_loss_values, metric_values

    def plot_loss_curve_and_mean_dice(self, epoch_loss_values, metric_values):
        """The reference's two-panel figure (VSparams.py:530-545), drawn with PIL, plus the same curves as CSV."""
        if self.rank != 0:
            return
        os.makedirs(self.figures_path, exist_ok=True)
        with open(os.path.join(self.figures_path, "epoch_average_loss_and_val_mean_dice.csv"), "w") as f:
            w = csv.writer(f)
            w.writerow(["epoch", "epoch_average_loss", "val_mean_dice"])
            for i, v in enumerate(epoch_loss_values):
                k = (i + 1) // self.val_interval - 1
                m = metric_values[k] if (i + 1) % self.val_interval == 0 and 0 <= k < len(metric_values) else ""
                w.writerow([i + 1, v, m])
        _plot_png(os.path.join(self.figures_path, "epoch_average_loss_and_val_mean_dice.png"),
                  [("Epoch Average Loss", [i + 1 for i in range(len(epoch_loss_values))], epoch_loss_values, "epoch"),
                   ("Val Mean Dice", [self.val_interval * (i + 1) for i in range(len(metric_values))], metric_values,
                    "epoch")])

    def load_trained_state_of_model(self, model):
        model.load_state_dict(torch.load(os.path.join(self.model_path, "best_metric_model.pth"),
                                         map_location=self.device))
        return model

    def run_inference(self, model, data_loader):
        """Sliding-window inference over the test set (reference VSparams.py:552-619).  On CUDA the finalise kernel
        emits the blended probabilities, the uint8 argmax mask and the Dice sums in one pass; the results of volume
        i are read back (one device->host copy of the mask + Dice) while volume i+1 is already running, so there
        is no per-image host sync on the launch path."""
        logger = self.logger
        logger.info("Running inference...")
        model.eval()
        dice_scores = np.zeros(len(data_loader))
        if self.model == "UNet2d5_spvPA":
            model_segmentation = lambda *args, **kwargs: model(*args, **kwargs)[0]  # noqa: E731
            model_segmentation.native_model = model  # fused sliding-window path on CUDA
        else:
            model_segmentation = model
        distributed = self.world_size > 1
        if distributed:
            from vs_seg_b200.parallel import sharded_sliding_window_inference
        from vs_seg_b200 import sliding_window as sw
        native = self.device.type == "cuda"

        def finish(i, data, mask, dice_score):
            """Host side of one volume: log, NIfTI export, figure (mask: uint8 argmax [1,1,X,Y,Z])."""
            mask = mask.cpu()
            dice_scores[i] = float(dice_score)
            logger.info(f"dice_score = {dice_scores[i]}")
            if self.rank == 0 and self.export_inferred_segmentations:
                logger.info("export to nifti...")
                meta = {k: (v[0] if isinstance(v, (list, tuple)) else v) for k, v in data["label_meta_dict"].items()}
                meta["affine"] = np.squeeze(np.asarray(meta["affine"]))
                meta["original_affine"] = np.squeeze(np.asarray(meta["original_affine"]))
                folder_name = os.path.basename(os.path.dirname(meta["filename_or_obj"]))
                saver = NiftiSaver(output_dir=os.path.join(self.results_folder_path, "inferred_segmentations_nifti",
                                                           folder_name), output_postfix="")
                saver.save(mask[0].to(torch.uint8), meta_data=meta)
            if self.rank == 0:
                label = torch.squeeze(data["label"][0, 0, :, :, :])
                slice_idx = self.get_center_of_mass_slice(label)
                os.makedirs(self.figures_path, exist_ok=True)
                _save_png(os.path.join(self.figures_path, "best_model_output_val" + str(i) + ".png"),
                          [data["image"][0, 0, :, :, slice_idx], data["label"][0, 0, :, :, slice_idx],
                           mask[0, 0, :, :, slice_idx]])

        pending = None
        with torch.no_grad():
            for i, data in enumerate(data_loader):
                logger.info("starting image {}".format(i))
                inputs = data["image"].to(self.device, non_blocking=True)
                roi = self.sliding_window_inferer_roi_size
                if native:
                    label = data["label"].to(torch.uint8).to(self.device, non_blocking=True)   # 1 byte per voxel
                    if distributed:
                        res = sharded_sliding_window_inference(inputs, roi, 1, model_segmentation, mode="gaussian",
                                                               label=label, return_mask=True)
                    else:
                        acc, cnt, lows, img = sw.sliding_window_accumulate(inputs, roi, model_segmentation,
                                                                           mode="gaussian", sw_batch_size=1)
                        res = sw.finalize(acc, cnt, lows, img, label=label, return_mask=True)
                    if res is None:  # only rank 0 holds the blended result
                        continue
                    outputs, mask, sums = res
                    dice_dev = sw.dice_from_sums(sums).mean()
                else:
                    if distributed:
                        outputs = sharded_sliding_window_inference(inputs, roi, 1, model_segmentation, mode="gaussian")
                        if outputs is None:
                            continue
                    else:
                        outputs = sliding_window_inference(inputs=inputs, roi_size=roi, sw_batch_size=1,
                                                           predictor=model_segmentation, mode="gaussian")
                    dice_dev = self.compute_dice_score(outputs, data["label"].to(self.device)).reshape(())
                    mask = torch.argmax(outputs, dim=1, keepdim=True)
                if pending is not None:   # volume i is queued: now read back volume i-1
                    finish(*pending)
                pending = (i, data, mask, dice_dev)
            if pending is not None:
                finish(*pending)

        if self.rank == 0:
            os.makedirs(self.figures_path, exist_ok=True)
            _hist_png(os.path.join(self.figures_path, "best_model_output_dice_score_histogram.png"), dice_scores,
                      np.arange(0, 1.01, 0.01))
        logger.info(f"all_dice_scores = {dice_scores}")
        logger.info(f"mean_dice_score = {dice_scores.mean()} +- {dice_scores.std()}")
        return dice_scores
