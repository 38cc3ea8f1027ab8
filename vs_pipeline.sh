#!/bin/bash
# train, then run inference with the same results folder (reference vs_pipeline.sh:5-6)
name=${1:-pipeline}
python3 VS_train.py --results_folder_name "$name" --dataset T1 "${@:2}" && python3 VS_inference.py --results_folder_name "$name" --dataset T1 "${@:2}"
