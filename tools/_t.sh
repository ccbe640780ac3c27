python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^(FAILED|ERROR|E  )|assert|Error" | head -30
