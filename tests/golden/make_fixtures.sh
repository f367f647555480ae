#!/bin/sh
# Regenerates tests/golden/ from the read-only reference checkout (byte copies of its test resources).
set -e
R=${1:-/root/reference}/test/resources/abalone
D=$(dirname "$0")
mkdir -p "$D/abalone"
cp "$R/models/libsvm_pickled/xgboost-model" "$D/abalone_xgboost-model.ubj"
cp "$R/data/train/abalone.train_0" "$R/data/train/abalone.train_1" "$R/data/validation/abalone.validation" "$D/abalone/"
