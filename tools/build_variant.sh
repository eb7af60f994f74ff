#!/bin/bash
# tools/build_variant.sh <name> <nvcc flags...>: builds variants/libthzdoe_<name>.so with extra -D flags (A/B experiments;
# select at run time with THZ_LIB=variants/libthzdoe_<name>.so).  The default library is rebuilt afterwards.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
THZ_NVCC_EXTRA="$*" python -m quantizationawarethzdoe_b200.build --force > /dev/null
cp quantizationawarethzdoe_b200/csrc/libthzdoe.so variants/libthzdoe_$name.so
echo built variants/libthzdoe_$name.so with "$*"
python -m quantizationawarethzdoe_b200.build --force > /dev/null
echo rebuilt the default library
