//! The reference CLI (src/main.rs) over the GPU engine: same flags, same stages and seeds, same
//! output lines.  `--plot/--output-path` are accepted; plotting belongs to the reference's
//! optional `debug` feature and is not part of this crate.
use clap::Parser;
use maxpro::anneal::anneal_lhd;
use maxpro::build_lhd;
use maxpro::enums::Metrics;
use maxpro::maximin_utils::maximin_criterion;
use maxpro::maxpro_utils::maxpro_criterion;
use maxpro::order::order_design;

#[derive(Parser)]
#[command(version, about = "Maximum projection / maximin Latin hypercube designs on a B200")]
struct Args {
    /// Number of samples
    #[arg(short, long)]
    samples: u64,
    /// Number of random candidate designs
    #[arg(short, long)]
    iterations: u64,
    /// Number of dimensions
    #[arg(short, long)]
    ndims: u64,
    /// Criterion
    #[arg(short, long, value_enum, default_value_t = Metrics::MaxPro)]
    metric: Metrics,
    /// Random seed (random if absent)
    #[arg(long)]
    seed: Option<u64>,
    #[arg(long, default_value_t = 100000)]
    anneal_iterations: u64,
    #[arg(long, default_value_t = 0.99)]
    anneal_cooling: f64,
    #[arg(long, default_value_t = 1.0)]
    anneal_t: f64,
    #[arg(short, long, default_value_t = false)]
    plot: bool,
    #[arg(short, long)]
    output_path: Option<String>,
    /// B200 engine only: devices the search and the chains are spread over (0 = all)
    #[arg(long, default_value_t = 1)]
    gpus: u32,
}

fn main() {
    let args = Args::parse();
    let seed = args.seed.unwrap_or_else(rand::random::<u64>);
    println!("Generating hypercube with seed {}", seed);
    assert!(args.samples > 0, "n_samples must be positive and nonzero");
    assert!(args.iterations > 0, "n_iterations must be positive and nonzero");
    assert!(args.ndims > 0, "n_dims must be positive and nonzero");
    assert!(args.anneal_iterations > 0, "annealing_iterations must be positive and nonzero");

    let (metric_fn, minimize) = match args.metric {
        Metrics::MaxPro => (maxpro_criterion as fn(&Vec<Vec<f64>>) -> f64, true),
        Metrics::MaxiMin => (maximin_criterion as fn(&Vec<Vec<f64>>) -> f64, false),
    };
    if args.gpus != 1 {
        maxpro::set_devices(&(0..args.gpus as i32).collect::<Vec<i32>>());
    }
    let lhd = build_lhd(args.samples, args.ndims, args.iterations, Some(args.metric.clone()), seed);
    let metric_value = metric_fn(&lhd);
    let swap_iterations = args.anneal_iterations / 5;
    let swapped = anneal_lhd(&lhd, swap_iterations, args.anneal_t, args.anneal_cooling, metric_fn, minimize, seed + 1, true);
    let swapped_metric = metric_fn(&swapped);
    let annealed = anneal_lhd(&swapped, args.anneal_iterations - swap_iterations, args.anneal_t, args.anneal_cooling,
                              metric_fn, minimize, seed + 2, false);
    let annealed_metric = metric_fn(&annealed);
    println!("{:?}", annealed);
    println!("Original metric: {metric_value}");
    println!("Swapped metric: {swapped_metric}");
    println!("Annealed metric: {annealed_metric}");
    let ordered = order_design(annealed, metric_fn, minimize);
    if args.plot {
        eprintln!("--plot needs the reference's `debug` feature (plotters); design ordered, {} points, nothing drawn{}",
                  ordered.len(), args.output_path.map_or(String::new(), |p| format!(" to {p}")));
    }
}
