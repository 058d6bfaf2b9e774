use clap::ValueEnum;

/// CLI spelling `max-pro` / `maxi-min` (clap kebab-case), as in the reference (src/enums.rs:3-7).
#[derive(ValueEnum, Clone, Debug, PartialEq, Eq)]
pub enum Metrics {
    MaxPro,
    MaxiMin,
}

impl Metrics {
    pub(crate) fn code(&self) -> std::os::raw::c_int {
        match self {
            Metrics::MaxPro => crate::ffi::METRIC_MAXPRO,
            Metrics::MaxiMin => crate::ffi::METRIC_MAXIMIN,
        }
    }
}
