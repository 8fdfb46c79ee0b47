"""Command-line front end of detsam2_b200.vos.vos_inference — same flags as the reference's tools/vos_inference.py
(:250-336) for the semi-supervised setting (masks of frame 0, or all available masks with --use_all_masks).

usage: python tools/vos_inference.py --sam2_cfg configs/sam2.1/sam2.1_hiera_l.yaml --sam2_checkpoint CKPT.pt \
           --base_video_dir DAVIS/JPEGImages/480p --input_mask_dir DAVIS/Annotations/480p --output_mask_dir out
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sam2_cfg", default="configs/sam2.1/sam2.1_hiera_b+.yaml")
    ap.add_argument("--sam2_checkpoint", default=None, help="reference-format checkpoint; seeded random weights if omitted")
    ap.add_argument("--base_video_dir", required=True)
    ap.add_argument("--input_mask_dir", required=True)
    ap.add_argument("--video_list_file", default=None)
    ap.add_argument("--output_mask_dir", required=True)
    ap.add_argument("--score_thresh", type=float, default=0.0)
    ap.add_argument("--use_all_masks", action="store_true")
    ap.add_argument("--per_obj_png_file", action="store_true")
    ap.add_argument("--apply_postprocessing", action="store_true")
    args = ap.parse_args()
    from detsam2_b200.build_sam import build_sam2_video_predictor
    from detsam2_b200.vos import vos_inference
    predictor = build_sam2_video_predictor(args.sam2_cfg, args.sam2_checkpoint, device="cuda",
                                           apply_postprocessing=args.apply_postprocessing)
    if args.video_list_file is not None:
        with open(args.video_list_file) as f:
            names = [v.strip() for v in f.readlines() if v.strip()]
    else:
        names = sorted(p for p in os.listdir(args.base_video_dir) if os.path.isdir(os.path.join(args.base_video_dir, p)))
    print(f"running VOS prediction on {len(names)} videos:\n{names}")
    for n, name in enumerate(names):
        print(f"\n{n + 1}/{len(names)} - running on {name}")
        vos_inference(predictor, args.base_video_dir, args.input_mask_dir, args.output_mask_dir, name,
                      score_thresh=args.score_thresh, use_all_masks=args.use_all_masks,
                      per_obj_png_file=args.per_obj_png_file)
    print(f"completed VOS prediction on {len(names)} videos -- output masks saved to {args.output_mask_dir}")


if __name__ == "__main__":
    main()
