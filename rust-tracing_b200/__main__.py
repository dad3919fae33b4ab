"""CLI with the reference's flags (main.rs:40-54): -l/--live, -s/--scene 0..8, -o/--output NAME (writes NAME.png),
plus the overrides the reference hard-codes per scene (--width, --spp, --depth) and seeds.

    python -m rust_tracing_b200 -s 6 -o cornell --spp 256
"""
import argparse
import sys
import time


def parse_args(argv=None):
    ap = argparse.ArgumentParser(prog="rust_tracing_b200", description="B200 path tracer with rust-tracing's scenes")
    ap.add_argument("-l", "--live", action="store_true", help="progressive passes (the reference's live preview, without the window)")
    ap.add_argument("-s", "--scene", type=int, default=0,
                    help="0:random balls, 1:two spheres, 2:earth, 3:perlin spheres, 4:quads, 5:simple light, 6:cornell box, 7:cornell smoke, 8:final scene")
    ap.add_argument("-o", "--output", default="output", help="name of the output file (.png is appended)")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--earth", default=None, help="path of assets/earth-large.jpg (a synthetic stand-in is used if absent)")
    ap.add_argument("--device", type=int, default=0)
    return ap.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    print(f"Args: {args}")                                   # main.rs:643
    from PIL import Image
    import rust_tracing_b200 as rt
    scene = args.scene if 0 <= args.scene <= 8 else 0       # main.rs:655
    ctx = rt.Context(args.device)
    earth = rt.load_earth(args.earth, ctx=ctx)[0] if scene in (2, 8) else None   # ImageTexture::new's decode, on the device
    t0 = time.time()
    s, cs = rt.builtin_scene(scene, image_width=args.width, samples_per_pixel=args.spp, max_depth=args.depth, earth=earth)
    cam = rt.Camera(cs)
    print(f"Building BVH: {time.time() - t0:.2f}s")          # main.rs:660 (scene + BVH build on the host)
    ds = ctx.upload(s)
    t0 = time.time()
    if args.live:
        from rust_tracing_b200.live import ProgressiveRender
        prog = ProgressiveRender(ctx, ds, cam, seed=args.seed, scene=s)
        frame = None
        for n, frame in prog.frames():
            if n % 16 == 0:
                print(f"rust-tracing [{cam.image_width}x{cam.image_height}, spp:{n}]")
        print(f"Render time: {time.time() - t0:.2f}s")
        if frame is not None:
            Image.fromarray(frame).save(f"{args.output}.png")
    else:
        rgb = ctx.render_rgb8(ds, cam, 0, cam.samples_per_pixel, args.seed)   # sums stay on the device; color_to_rgb runs there
        dt = time.time() - t0
        h, w = cam.shape
        print(f"Render time: {dt:.2f}s ({h * w * cam.samples_per_pixel / dt / 1e6:.1f} Mpaths/s)")   # renderer.rs:51
        t0 = time.time()
        Image.fromarray(rgb).save(f"{args.output}.png")
        print(f"PNG encoding: {time.time() - t0:.2f}s")      # renderer.rs:73
    return 0


if __name__ == "__main__":
    sys.exit(main())
